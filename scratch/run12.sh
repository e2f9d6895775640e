out=gpurun_out
for w in cfg2 cfg3; do
  ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 4 -c 1 -f -o $out/prof_r1c_${w}_1frame \
      python tests/analysis/kbench.py $w:1 --steps 5 > $out/ncu_single_$w.log 2>&1
done
