# class-split launches: parity with the product build, then timings; experiment build for class-only runs
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
$K --tag "split (31K / 50K)"
PB_CLASS_SPLIT=0 $K --tag "one launch"
PB_ONE_KIB=26 $K --tag "split one=26K"
PB_ONE_KIB=20 $K --tag "split one=20K"
PB_REST_KIB=40 $K --tag "split rest=40K"
PB_L2_AHEAD=1 $K --tag "split l2_ahead=1"
PB_RASTER_BAND=8 $K --tag "split band=8"
PB_RASTER_BAND=30 $K --tag "split band=30"
export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
PB_DEBUG_MODE=32 $K --tag "exp split: one + blend"
PB_DEBUG_MODE=64 $K --tag "exp split: one + both"
PB_DEBUG_MODE=96 $K --tag "exp split: one only"
PB_DEBUG_MODE=32 PB_L2_AHEAD=2 $K --tag "exp split: one + blend, l2 ahead 2"
PB_DEBUG_MODE=32 PB_RASTER_BAND=6 $K --tag "exp split: one + blend, band 6"
