out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mid_size or batch_equals or small_matrix_against_golden or fast_path" 2>&1 | tail -3 > $out/wgt_test.log
cat $out/wgt_test.log
timeout 120 python tests/analysis/kbench.py cfg5:16 cfg5:32 cfg4:1 T:16 --tag "lean wgt" > $out/kbench_wgt.log 2>&1
cat $out/kbench_wgt.log
