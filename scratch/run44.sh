timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "blend_band" 2>&1 | tail -2
export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
K="timeout 100 python tests/analysis/kbench.py"
$K cfg5:16 T:16 --tag "loads evict_last (default)"
PB_DEBUG_MODE=8 $K cfg5:16 T:16 --tag "loads evict_first"
PB_DEBUG_MODE=128 $K cfg5:16 T:16 --tag "loads evict_normal"
