python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="python tests/analysis/kbench.py T:16 cfg5:16 T:1 cfg4:1 cfg1:1"
$K --tag "syncfree default"
for n in 3 4; do PB_STAGE_BUFFERS=$n python tests/analysis/kbench.py T:16 cfg5:16 --tag "syncfree buffers $n"; done
