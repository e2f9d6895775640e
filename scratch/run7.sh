for wl in cfg5 T; do
ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 3 -c 1 -f -o gpurun_out/prof_r1h_$wl python tests/analysis/kbench.py $wl:16 --steps 2 --warmup 1 > gpurun_out/ncu_r1h_$wl.log 2>&1
done
tail -3 gpurun_out/ncu_r1h_T.log
