"""Command line of the B200-native ProNeRF render path.

Mirrors ``pronerf/cli.py`` of the reference (sub-commands ``infer`` / ``eval`` with ``--config``,
``--checkpoint``, ``--render-test``, ``--use-trt``, ``--max-images`` and ``--`` pass-through, cli.py:95-102,
195-210).  ``infer`` / ``eval`` run ``pronerf_b200.render.train`` (the reference's infer driver, trt.py:699);
the training and TensorRT-export sub-commands exist so scripts get a clear message instead of a usage error.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path

REPO_ROOT = Path(__file__).resolve().parents[1]
DEFAULT_TRT_CONFIG = REPO_ROOT / "configs/llff/fern/fern_b200.txt"

OUT_OF_SCOPE = ("{cmd}: training and TensorRT export are outside the render hot path that pronerf_b200 implements "
                "(SURVEY.md section 8); use the upstream repository for them.")


def _extra_args(args) -> list:
    extra = list(args.extra)
    if extra and extra[0] == "--":
        extra = extra[1:]
    return extra


def build_argv(args) -> list:
    """The argv the reference CLI would hand to its infer script (cli.py:95-102)."""
    cfg = Path(args.config)
    if not cfg.is_absolute():
        cfg = REPO_ROOT / cfg
    argv = ["--config", str(cfg)]
    if args.checkpoint is not None:
        argv += ["--ft_path", str(args.checkpoint)]
    if args.render_test:
        argv.append("--render_test")
    if args.use_trt:
        argv.append("--use_trt")
    if args.max_images is not None:
        argv += ["--max_images", str(args.max_images)]
    return argv + _extra_args(args)


def infer(args):
    from pronerf_b200.render import train
    return train(build_argv(args))


def eval_model(args):
    args.render_test = True
    return infer(args)


def out_of_scope(args):
    raise SystemExit(OUT_OF_SCOPE.format(cmd=args.command))


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="python -m pronerf.cli",
                                     description="B200-native ProNeRF render path (LLFF fern-shaped inference).")
    sub = parser.add_subparsers(dest="command", required=True)

    def passthrough(p):
        p.add_argument("extra", nargs=argparse.REMAINDER,
                       help="additional arguments forwarded to the infer driver; prefix with --")

    p = sub.add_parser("infer", help="render held-out/test views")
    p.add_argument("--config", default=str(DEFAULT_TRT_CONFIG))
    p.add_argument("--checkpoint", default=None)
    p.add_argument("--render-test", action="store_true", dest="render_test")
    p.add_argument("--use-trt", action="store_true", dest="use_trt")
    p.add_argument("--max-images", type=int, default=None, dest="max_images")
    passthrough(p)
    p.set_defaults(func=infer)

    p = sub.add_parser("eval", help="render test split through the inference path")
    p.add_argument("--config", default=str(DEFAULT_TRT_CONFIG))
    p.add_argument("--checkpoint", default=None)
    p.add_argument("--use-trt", action="store_true", dest="use_trt")
    p.add_argument("--max-images", type=int, default=None, dest="max_images")
    passthrough(p)
    p.set_defaults(func=eval_model, render_test=True)

    for name in ("train-stage1", "train-stage2", "export-trt"):
        p = sub.add_parser(name, help="not part of pronerf_b200 (see upstream)")
        passthrough(p)
        p.set_defaults(func=out_of_scope)
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    os.chdir(REPO_ROOT)
    return args.func(args)


if __name__ == "__main__":
    main()
