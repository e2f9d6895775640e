timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 120 python tests/analysis/kbench.py"
$K cfg5:16 T:16 --tag "defaults (one 21K x4, rest 50K x2)"
PB_REST_KIB=90 $K cfg5:16 --tag "rest 90K (1 CTA/SM)"
PB_REST_KIB=44 $K cfg5:16 --tag "rest 44K"
PB_RASTER_BAND=6 $K cfg5:16 --tag "band 6"
PB_RASTER_BAND=60 $K cfg5:16 --tag "band 60"
$K cfg5:8 cfg5:32 cfg5:4 --tag "other batch sizes"
