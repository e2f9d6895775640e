out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/direct_test.log
cat $out/direct_test.log
timeout 120 python tests/analysis/kbench.py cfg2:1 cfg3:1 --tag direct > $out/kbench_direct.log 2>&1
PB_DIRECT=0 timeout 120 python tests/analysis/kbench.py cfg2:1 cfg3:1 --tag tiled >> $out/kbench_direct.log 2>&1
cat $out/kbench_direct.log
