out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mid_size" 2>&1 | tail -3 > $out/sep1_test.log
cat $out/sep1_test.log
K="timeout 120 python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1"
( $K --tag rot
PB_SEP1_WAVES=2 $K --tag "rot waves2"
PB_SEP1_WAVES=3 $K --tag "rot waves3"
) > $out/kbench_sep1_rot.log 2>&1
cat $out/kbench_sep1_rot.log
