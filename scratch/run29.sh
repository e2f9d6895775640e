# re-entry check of HEAD: GPU tests, smoke, default bench, kernel timings
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/s6_pytest.log 2>&1; tail -3 gpurun_out/s6_pytest.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 400 python bench.py ) > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; tail -c 600 gpurun_out/s6_bench.json
timeout 200 python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1 cfg2:1 cfg3:1 T:16 cfg5:16 --tag s6 2>&1 | tee gpurun_out/s6_kbench.log
