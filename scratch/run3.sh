python -m pytest tests -m gpu -x -q 2>&1 | tail -2
K="python tests/analysis/kbench.py T:16 cfg5:16"
for d in 0 1 2 3 4 6 8; do PB_L2_AHEAD=$d $K --tag "l2_ahead $d"; done
