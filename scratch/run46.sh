timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
K="timeout 100 python tests/analysis/kbench.py"
$K cfg4:1 --tag "sep1 two grids"
PB_CLASS_SPLIT=0 $K cfg4:1 --tag "sep1 one grid"
PB_SEP1_ONE_KIB=20 $K cfg4:1 --tag "sep1 two grids, one=20K"
PB_SEP1_ONE_KIB=12 $K cfg4:1 --tag "sep1 two grids, one=12K"
PB_SEP1_KIB=40 $K cfg4:1 --tag "sep1 two grids, rest=40K"
