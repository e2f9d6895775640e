export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
$K --tag "all"
PB_DEBUG_MODE=48 $K --tag "blend only"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=48 $K --tag "blend, 2 CTA/SM 2x48K"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=16 $K --tag "both+blend, 2 CTA/SM 2x48K"
