for m in 1:boxes 0:rects; do
  mode=${m%%:*}; name=${m##*:}
  PB_BOXES=$mode ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 4 -c 1 -f -o gpurun_out/T16_$name \
     python tests/analysis/kbench.py T:16 --steps 5 > gpurun_out/ncu_T16_$name.log 2>&1
  (python profiles/ncu_summary.py gpurun_out/T16_$name.ncu-rep --stalls; python profiles/ncu_sass_hot.py gpurun_out/T16_$name.ncu-rep 1.0) > gpurun_out/T16_$name.txt 2>&1
done
rm -f gpurun_out/T16_rects.ncu-rep
