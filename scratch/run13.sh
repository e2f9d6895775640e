out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mid_size or round_trip or batch_equals or small_matrix_against_golden" -s 2>&1 | tail -12 > $out/sep1_test.log
cat $out/sep1_test.log
timeout 120 python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1 --tag sep1 > $out/kbench_sep1.log 2>&1
PB_SEP1=0 timeout 120 python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1 --tag old >> $out/kbench_sep1.log 2>&1
cat $out/kbench_sep1.log
