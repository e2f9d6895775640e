# cfg5 x16 by tile class (experiment build: PB_DEBUG_MODE bits 16 / 32 / 64 leave a class out)
export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
$K --tag "all"
PB_DEBUG_MODE=96 $K --tag "single only"
PB_DEBUG_MODE=80 $K --tag "both-unit only"
PB_DEBUG_MODE=48 $K --tag "blend only"
PB_DEBUG_MODE=16 $K --tag "both+blend"
PB_DEBUG_MODE=64 $K --tag "no blend"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=80 $K --tag "both-unit, 2 CTA/SM 48K"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=48 $K --tag "blend, 2 CTA/SM 48K"
PB_TWO_BUF_LIMIT_KIB=113 PB_STAGE_KIB=48 PB_DEBUG_MODE=96 $K --tag "single, 2 CTA/SM 48K"
