#!/bin/bash
# Tuning aid: time the three tensor-core MLP kernels for several slot skews (PN_TC_SHIFT) and split publishing.
for sh in 0 2 3 4 5; do PN_TC_SHIFT=$sh python scripts/tc_sweep.py 2>&1 | tail -1; done
PN_TC_SHIFT=0 PN_TC_SPLIT=1 python scripts/tc_sweep.py 2>&1 | tail -1
PN_TC_SHIFT=4 PN_TC_SPLIT=1 python scripts/tc_sweep.py 2>&1 | tail -1
