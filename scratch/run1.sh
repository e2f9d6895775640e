python -m pytest tests -m gpu -x -q 2>&1 | tail -3
K="python tests/analysis/kbench.py cfg5:16 T:16 cfg1:1 cfg4:1"
PB_RASTER_BAND=0 PB_PITCH_EVEN=1 $K --tag "band0 even"
PB_RASTER_BAND=0 $K --tag "band0 odd"
PB_RASTER_BAND=4 $K --tag "band4 odd"
PB_RASTER_BAND=8 $K --tag "band8 odd"
PB_RASTER_BAND=16 $K --tag "band16 odd"
PB_RASTER_BAND=30 $K --tag "band30 odd"
PB_RASTER_BAND=60 $K --tag "band60 odd"
