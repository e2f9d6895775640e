# single-frame captures: what bounds one frame per launch
out=gpurun_out
python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1 cfg2:1 cfg3:1 T:4 cfg5:4 --tag single > $out/kbench_single.log 2>&1
for w in T cfg4 cfg2 cfg3; do
  ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 4 -c 1 -f -o $out/prof_r1b_${w}_1frame \
      python tests/analysis/kbench.py $w:1 --steps 5 > $out/ncu_single_$w.log 2>&1
done
cat $out/kbench_single.log
