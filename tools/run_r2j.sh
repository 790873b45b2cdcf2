set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -5 gpurun_out/r2j_pytest.log
GMSB_TC_TRACE=1 timeout 300 python tools/e2e_trace.py --scale 24 --reps 3 --shard-parts 8 > gpurun_out/r2j_e2e_trace.jsonl 2> gpurun_out/r2j_e2e_trace.err; tail -4 gpurun_out/r2j_e2e_trace.jsonl; grep "trace" gpurun_out/r2j_e2e_trace.err | tail -40
timeout 400 python bench.py --steps 10 --warmup 3 --kclique '' --no-cpu-baseline > gpurun_out/r2j_bench_1gpu_quick.json 2> gpurun_out/r2j_bench_1gpu_quick.err; tail -3 gpurun_out/r2j_bench_1gpu_quick.err; cut -c1-300 gpurun_out/r2j_bench_1gpu_quick.json
