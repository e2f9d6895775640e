python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tests/analysis/kbench.py T:16 cfg5:16 T:1 cfg4:1 cfg1:1 cfg2:1 cfg3:1 --tag "boxes"
for k in 12 16 20 24; do PB_STAGE_KIB=$k python tests/analysis/kbench.py cfg5:16 --tag "stage KiB $k"; done
for k in 6 8 10 12; do PB_STAGE_KIB=$k python tests/analysis/kbench.py T:16 --tag "stage KiB $k"; done
