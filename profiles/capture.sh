#!/bin/bash
# Round evidence, run on the GPU box:  gpurun -- 'bash profiles/capture.sh r1'
# 1. plain bench run (must exit 0)  2. ncu launch list of the same command
# 3. ncu --set full capture of the dominant kernel of the default workload (and of the 8K target)
tag=${1:-r1}
out=gpurun_out
small="--steps 5 --warmup 3 --no-cpu-baseline --no-also --e2e-steps 1"
python bench.py $small > $out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_${tag}.csv \
    python bench.py $small > $out/ncu_launches_${tag}.log 2>&1
python bench.py $small > $out/plain_${tag}b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 4 -c 1 -f -o $out/prof_${tag}_cfg5 \
    python bench.py $small > $out/ncu_full_${tag}.log 2>&1
python bench.py --workload T $small > $out/plain_${tag}_T.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 4 -c 1 -f -o $out/prof_${tag}_T \
    python bench.py --workload T $small > $out/ncu_full_${tag}_T.log 2>&1
tail -2 $out/ncu_full_${tag}_T.log
