timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
K="timeout 100 python tests/analysis/kbench.py"
$K cfg5:16 cfg4:1 --tag "two grids, concurrent"
PB_CONCURRENT=0 $K cfg5:16 cfg4:1 --tag "two grids, in sequence"
$K cfg5:16 cfg4:1 --tag "two grids, concurrent"
PB_CONCURRENT=0 $K cfg5:16 cfg4:1 --tag "two grids, in sequence"
