out=gpurun_out
for w in T cfg4; do
  ncu --set full --clock-control none --import-source on -k regex:remap_sep1 -s 4 -c 1 -f -o $out/prof_r1e_${w}_1frame \
      python tests/analysis/kbench.py $w:1 --steps 5 > $out/ncu_sep1_$w.log 2>&1
done
for g in 296 444 592 740 888 1480; do PB_SEP1_GRID=$g timeout 120 python tests/analysis/kbench.py T:1 cfg4:1 --tag "grid $g" ; done > $out/kbench_sep1_grid.log 2>&1
cat $out/kbench_sep1_grid.log
