out=gpurun_out
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
( $K --tag "64reg base"
for k in 18 20 22 44 48; do PB_STAGE_KIB=$k $K --tag "64reg stage $k"; done
) > $out/kbench_cfg5_64reg.log 2>&1
cat $out/kbench_cfg5_64reg.log
