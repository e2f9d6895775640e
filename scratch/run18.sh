out=gpurun_out
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
( $K --tag base
for k in 36 40 44 48; do PB_TWO_BUF_LIMIT_KIB=120 PB_STAGE_KIB=$k $K --tag "2buf stage $k"; done
PB_TWO_BUF_LIMIT_KIB=120 $K --tag "2buf tuned stage"
) > $out/kbench_cfg5_2buf.log 2>&1
cat $out/kbench_cfg5_2buf.log
