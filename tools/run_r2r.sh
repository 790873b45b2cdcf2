set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log; tail -4 gpurun_out/r2r_pytest.log
timeout 600 python bench.py > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; tail -3 gpurun_out/bench_r2r.err; cut -c1-300 gpurun_out/bench_r2r.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2r.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-clocks --kclique '' > gpurun_out/bench_under_ncu_r2r.json 2> gpurun_out/bench_under_ncu_r2r.err
wc -l gpurun_out/launches_r2r.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_plan_scatter|k_plan_count|k_emit_rows|k_move_rows|k_relabel_count|k_sort_mid|k_classify" -c 14 -o gpurun_out/r2r_prof_prep python tools/tc_sweep.py --scale 24 --reps 1 --configs '[{"variant":"auto"}]' > gpurun_out/r2r_prof_prep.log 2>&1
ls -la gpurun_out/r2r_prof_prep.ncu-rep
