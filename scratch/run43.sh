# producer warp: parity first (a hang is cut by timeout), then timings with and without
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 100 python tests/analysis/kbench.py"
$K cfg5:16 T:16 --tag "producer warp"
PB_PRODUCER=0 $K cfg5:16 T:16 --tag "no producer"
$K T:1 cfg4:1 cfg5:4 --tag "producer warp"
