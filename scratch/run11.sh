out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fast_path or full_size" -s 2>&1 | tail -15 > $out/fast_test.log
cat $out/fast_test.log
python tests/analysis/kbench.py cfg2:1 cfg3:1 --tag fast > $out/kbench_fast.log 2>&1
PB_EXACT_CHAIN=1 python tests/analysis/kbench.py cfg2:1 cfg3:1 --tag exact >> $out/kbench_fast.log 2>&1
cat $out/kbench_fast.log
