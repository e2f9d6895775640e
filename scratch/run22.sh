out=gpurun_out
for w in cfg2 cfg3; do
  ncu --set full --clock-control none --import-source on -k regex:remap_direct -s 4 -c 1 -f -o $out/prof_r1g_${w}_1frame \
      python tests/analysis/kbench.py $w:1 --steps 5 > $out/ncu_single_$w.log 2>&1
  (python profiles/ncu_summary.py $out/prof_r1g_${w}_1frame.ncu-rep --stalls; python profiles/ncu_sass_segments.py $out/prof_r1g_${w}_1frame.ncu-rep) > $out/r1g_ncu_full_${w}_1frame.txt 2>&1
done
rm -f $out/prof_r1g_cfg3_1frame.ncu-rep
