K="timeout 120 python tests/analysis/kbench.py"
PB_BOXES=0 $K cfg5:16 T:16 --tag "rectangles"
PB_BOX_WIDTHS=0 $K cfg5:16 T:16 --tag "boxes, own widths"
PB_BOX_WIDTHS=1 $K cfg5:16 T:16 --tag "boxes, 1 width"
PB_BOX_WIDTHS=2 $K cfg5:16 T:16 --tag "boxes, 2 widths"
PB_BOX_WIDTHS=2 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
