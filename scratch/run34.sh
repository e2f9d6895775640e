export PB_REMAP_LIB=$PWD/photonbend_b200/libpbremap_exp.so
for m in 48:blend; do
  mode=${m%%:*}; name=${m##*:}
  PB_DEBUG_MODE=$mode ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 4 -c 1 -f -o gpurun_out/cls2_$name \
     python tests/analysis/kbench.py cfg5:16 --steps 5 > gpurun_out/ncu_cls2_$name.log 2>&1
  (python profiles/ncu_summary.py gpurun_out/cls2_$name.ncu-rep --stalls; python profiles/ncu_sass_hot.py gpurun_out/cls2_$name.ncu-rep 0.7) > gpurun_out/cls2_$name.txt 2>&1
done
