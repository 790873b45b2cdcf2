#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same command, ncu --set full of the triangle
# kernels (auto schedule, and the forced all-merge / all-gallop schedules that justify the per-edge choice),
# kernel-choice sweep.  Outputs land in gpurun_out/ ; tools/summarise_profiles.py turns them into profiles/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -2 gpurun_out/bench_${TAG}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-clocks --kclique '' \
    > gpurun_out/bench_under_ncu_${TAG}.json 2> gpurun_out/bench_under_ncu_${TAG}.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_bitmap|k_tc_merge|k_tc_gallop" -s 4 -c 4 \
    -o gpurun_out/prof_tc_${TAG} python tools/tc_sweep.py --scale 24 --reps 2 --configs '[{"variant":"auto"}]' \
    > gpurun_out/prof_tc_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_tc_merge|k_tc_gallop" -c 2 \
    -o gpurun_out/prof_forced_${TAG} python tools/tc_sweep.py --scale 22 --reps 1 \
    --configs '[{"variant":"merge"},{"variant":"gallop"}]' > gpurun_out/prof_forced_${TAG}.log 2>&1
timeout 900 python tools/tc_sweep.py --scale 24 --reps 3 \
    --configs '[{"variant":"auto"},{"variant":"bitmap"},{"variant":"merge"},{"variant":"gallop"}]' \
    > gpurun_out/sweep_variants_${TAG}.jsonl 2> gpurun_out/sweep_variants_${TAG}.err
cut -c1-300 gpurun_out/bench_${TAG}.json
