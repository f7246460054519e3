"""``python -m pronerf.cli`` entry point (same package name and sub-commands as the reference's pronerf/cli.py),
dispatching to the B200-native implementation in ``pronerf_b200``."""
__version__ = "0.1.0"
