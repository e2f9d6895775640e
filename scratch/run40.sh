timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 120 python tests/analysis/kbench.py"
$K cfg5:16 T:16 --tag "spread issue, boxes"
PB_BOXES=0 $K cfg5:16 T:16 --tag "spread issue, rectangles"
PB_BOX_WIDTHS=2 $K cfg5:16 T:16 --tag "spread issue, boxes 2 widths"
PB_ONE_BYTES=21504 $K cfg5:16 --tag "boxes one=21K"
PB_ONE_BYTES=16384 $K cfg5:16 --tag "boxes one=16K"
