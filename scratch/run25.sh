timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mid_size or tiny or row_bands or batch_equals" 2>&1 | tail -3
timeout 120 python tests/analysis/kbench.py cfg4:1 --tag "items"
PB_SEP1_ITEMS=0 timeout 120 python tests/analysis/kbench.py cfg4:1 --tag "tile-staged"
for k in 20 24 28 32; do PB_STAGE_KIB=$k timeout 120 python tests/analysis/kbench.py cfg4:1 --tag "items stage $k"; done
