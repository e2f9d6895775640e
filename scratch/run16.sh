out=gpurun_out
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
( $K --tag base
PB_LEAN_MIN_GROUPS=2 $K --tag "lean>=2"
PB_LEAN_MIN_GROUPS=3 $K --tag "lean>=3"
for k in 16 20 24 28 32 36 44; do PB_STAGE_KIB=$k $K --tag "stage $k"; done
for k in 20 24 36; do PB_LEAN_MIN_GROUPS=2 PB_STAGE_KIB=$k $K --tag "stage $k lean>=2"; done
for b in 4 8 30 60; do PB_RASTER_BAND=$b $K --tag "band $b"; done
) > $out/kbench_cfg5_knobs.log 2>&1
cat $out/kbench_cfg5_knobs.log
