# staircase of boxes (per-box extents) in the lean loop: parity, then timings with and without
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 120 python tests/analysis/kbench.py"
$K cfg5:16 T:16 --tag "boxes"
PB_BOXES=0 $K cfg5:16 T:16 --tag "rectangles"
PB_ONE_BYTES=21504 $K cfg5:16 --tag "boxes one=21K"
PB_ONE_BYTES=16384 $K cfg5:16 --tag "boxes one=16K"
PB_ONE_BYTES=12288 $K cfg5:16 --tag "boxes one=12K"
PB_REST_KIB=40 $K cfg5:16 --tag "boxes rest=40K"
PB_STAGE_KIB=12 $K T:16 --tag "boxes T stage 12K"
PB_STAGE_KIB=20 $K T:16 --tag "boxes T stage 20K"
