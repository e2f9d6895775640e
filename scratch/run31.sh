# weighted lean loop + fixed-point blend: parity, then cfg5 x16 by tile class
export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
$K --tag "all"
PB_DEBUG_MODE=48 $K --tag "blend only"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=48 $K --tag "blend, 2 CTA/SM 2x48K"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=16 $K --tag "both+blend, 2 CTA/SM 2x48K"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=40 PB_DEBUG_MODE=16 $K --tag "both+blend, 2 CTA/SM 2x40K"
PB_TWO_BUF_LIMIT_KIB=226 PB_STAGE_KIB=96 PB_DEBUG_MODE=16 $K --tag "both+blend, 1 CTA/SM 2x96K"
PB_TWO_BUF_LIMIT_KIB=75 PB_STAGE_KIB=30 PB_DEBUG_MODE=96 $K --tag "single, 3 CTA/SM 2x30K"
PB_TWO_BUF_LIMIT_KIB=75 PB_STAGE_KIB=20 PB_DEBUG_MODE=96 $K --tag "single, 3 CTA/SM 2x20K"
PB_DEBUG_MODE=112 $K --tag "nothing (skip cost)"
