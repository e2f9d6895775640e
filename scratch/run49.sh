timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
K="timeout 100 python tests/analysis/kbench.py"
$K cfg5:16 cfg4:1 --tag "default (rest 50K, one 21K)"
PB_REST_KIB=36 $K cfg5:16 --tag "rest 36K"
PB_REST_KIB=44 $K cfg5:16 --tag "rest 44K"
PB_REST_KIB=36 PB_ONE_BYTES=16384 $K cfg5:16 --tag "rest 36K one 16K"
PB_REST_KIB=30 PB_ONE_BYTES=16384 $K cfg5:16 --tag "rest 30K one 16K"
