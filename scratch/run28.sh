timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mid_size or tiny or row_bands" 2>&1 | tail -2
PB_SEP1_CHUNKED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mid_size or tiny or row_bands" 2>&1 | tail -2
K="timeout 120 python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1"
$K --tag "interleaved"
PB_SEP1_CHUNKED=1 $K --tag "chunked"
