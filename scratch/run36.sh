timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
K="timeout 120 python tests/analysis/kbench.py cfg5:16"
$K --tag "split one=31K (3/SM)"
PB_ONE_BYTES=21504 $K --tag "split one=21K (4/SM)"
PB_ONE_BYTES=16384 $K --tag "split one=16K (4/SM)"
export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
PB_ONE_BYTES=21504 $K --tag "46 regs one=21K (4/SM)"
PB_ONE_BYTES=16384 $K --tag "46 regs one=16K (5/SM)"
PB_ONE_BYTES=16384 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "batch or frames or cfg5" 2>&1 | tail -2
PB_ONE_BYTES=16384 ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 8 -c 2 -f -o gpurun_out/split_cfg5 \
     python tests/analysis/kbench.py cfg5:16 --steps 5 > gpurun_out/ncu_split.log 2>&1
(python profiles/ncu_summary.py gpurun_out/split_cfg5.ncu-rep --stalls; python profiles/ncu_sass_hot.py gpurun_out/split_cfg5.ncu-rep 1.0) > gpurun_out/split_cfg5.txt 2>&1
