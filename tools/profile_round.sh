#!/bin/bash
# Regenerates the scratch inputs of profiles/ on a GPU box (run through gpurun, one GPU):
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2'
# then, back in the container:  python tools/summarize_profiles.py r2 gpurun_out/r2_launches.csv gpurun_out/prof_r2.ncu-rep 32
# Numbers printed by a run under ncu are never bench values; the bench lines come from the plain runs below.
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python tools/op_sweep.py gpurun_out/${tag}_op_sweep.csv >> gpurun_out/${tag}_bench.err 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fft2_|normal_' -s 5 -c 5 -f -o gpurun_out/prof_${tag} \
    python tools/prof_target.py 32 15 10 200 200 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -1 gpurun_out/${tag}_bench.json
tail -1 gpurun_out/${tag}_bench_reference.json
