#!/bin/bash
# Round evidence, run on the GPU box:  gpurun -- 'bash profiles/capture.sh r2'
# 1. plain default bench run (must exit 0) + the reference arm
# 2. ncu launch list of the bench command
# 3. ncu --set full of the dominant kernels: the default workload (cfg5 x16: two grids), the 8K target x16,
#    and the single-frame launches of T, cfg1, cfg4, cfg2, cfg3
# 4. summaries (.txt) and traffic_<tag>.json made here (ncu is on the box); copy them to profiles/
tag=${1:-r2}
out=gpurun_out
small="--steps 5 --warmup 3 --no-cpu-baseline --no-also --e2e-steps 1 --sustain-seconds 0 --no-parity --no-compressed"
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref_${tag}.json 2> $out/bench_ref_${tag}.err
python bench.py $small > $out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_${tag}.csv \
    python bench.py $small > $out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:remap_ -s 8 -c 2 -f -o $out/prof_${tag}_cfg5_16frames \
    python bench.py $small > $out/ncu_full_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:remap_ -s 4 -c 1 -f -o $out/prof_${tag}_T_16frames \
    python bench.py --workload T $small > $out/ncu_full_${tag}_T.log 2>&1
python tests/analysis/kbench.py T:1 cfg1:1 cfg4:1 cfg2:1 cfg3:1 T:16 cfg5:16 --tag $tag > $out/kbench_${tag}.log 2>&1
for w in T cfg1 cfg4 cfg2 cfg3; do
  n=1; [ $w = cfg4 ] && n=2   # a double-fisheye source is two grids per call (tile classes)
  ncu --set full --clock-control none --import-source on -k regex:remap_ -s 4 -c $n -f -o $out/prof_${tag}_${w}_1frame \
      python tests/analysis/kbench.py $w:1 --steps 5 > $out/ncu_single_${tag}_$w.log 2>&1
done
for r in cfg5_16frames T_16frames; do
  (python profiles/ncu_summary.py $out/prof_${tag}_$r.ncu-rep --stalls; python profiles/ncu_sass_hot.py $out/prof_${tag}_$r.ncu-rep 2.0) > $out/${tag}_ncu_full_$r.txt 2>&1
done
for w in T cfg1 cfg4 cfg2 cfg3; do
  (python profiles/ncu_summary.py $out/prof_${tag}_${w}_1frame.ncu-rep --stalls; python profiles/ncu_sass_segments.py $out/prof_${tag}_${w}_1frame.ncu-rep) > $out/${tag}_ncu_full_${w}_1frame.txt 2>&1
done
python profiles/ncu_traffic.py --out $out/traffic_${tag}.json \
    cfg5:16=$out/prof_${tag}_cfg5_16frames.ncu-rep T:16=$out/prof_${tag}_T_16frames.ncu-rep \
    T:1=$out/prof_${tag}_T_1frame.ncu-rep cfg1:1=$out/prof_${tag}_cfg1_1frame.ncu-rep cfg4:1=$out/prof_${tag}_cfg4_1frame.ncu-rep \
    cfg2:1=$out/prof_${tag}_cfg2_1frame.ncu-rep cfg3:1=$out/prof_${tag}_cfg3_1frame.ncu-rep > $out/traffic_${tag}.log 2>&1
# only one report travels back (64 MiB limit on gpurun_out/)
rm -f $out/prof_${tag}_T_16frames.ncu-rep $out/prof_${tag}_*_1frame.ncu-rep
cat $out/kbench_${tag}.log
