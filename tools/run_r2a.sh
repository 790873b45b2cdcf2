set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
GMSB_TC_TRACE=1 timeout 300 python tools/e2e_trace.py --scale 24 --reps 3 > gpurun_out/r2a_e2e_trace.jsonl 2> gpurun_out/r2a_e2e_trace.err
timeout 400 python bench.py --steps 10 --warmup 3 --kclique '' --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
GMSB_KCLIQUE_TRACE=1 timeout 500 python tools/kc_orient_ab.py --scale 22 --ks 5,6 > gpurun_out/r2a_kc_orient.jsonl 2> gpurun_out/r2a_kc_orient.err
tail -3 gpurun_out/r2a_e2e_trace.jsonl; cat gpurun_out/r2a_bench.json | cut -c1-600; cat gpurun_out/r2a_kc_orient.jsonl
