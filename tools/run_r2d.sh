set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
GMSB_TC_TRACE=1 timeout 300 python tools/e2e_trace.py --scale 24 --reps 3 --orient 1 > gpurun_out/r2d_e2e_trace.jsonl 2> gpurun_out/r2d_e2e_trace.err
timeout 300 python tools/e2e_trace.py --scale 24 --reps 3 --orient 0 >> gpurun_out/r2d_e2e_trace.jsonl 2>> gpurun_out/r2d_e2e_trace0.err
cat gpurun_out/r2d_e2e_trace.jsonl
timeout 400 python bench.py --steps 10 --warmup 3 --kclique '' --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cut -c1-300 gpurun_out/r2d_bench.json
GMSB_KCLIQUE_TRACE=1 timeout 400 python tools/kc_prof.py 20 7 > gpurun_out/r2d_kc_s20_k7.log 2>&1
cat gpurun_out/r2d_kc_s20_k7.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_bitmap2" -s 3 -c 3 \
    -o gpurun_out/r2d_prof_bitmap python tools/tc_sweep.py --scale 24 --reps 2 --configs '[{"variant":"auto"}]' > gpurun_out/r2d_prof_bitmap.log 2>&1
ls -la gpurun_out/r2d_prof_bitmap.ncu-rep
