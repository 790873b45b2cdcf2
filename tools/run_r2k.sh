set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -5 gpurun_out/r2k_pytest.log
timeout 300 python tools/partition_balance.py --scale 24 --parts 1,2,4,8 > gpurun_out/r2k_partition_balance.jsonl 2> gpurun_out/r2k_partition_balance.err; cat gpurun_out/r2k_partition_balance.jsonl; tail -3 gpurun_out/r2k_partition_balance.err
GMSB_TC_TRACE=1 timeout 300 python tools/e2e_trace.py --scale 24 --reps 3 --shard-parts 8 > gpurun_out/r2k_e2e_trace.jsonl 2> gpurun_out/r2k_e2e_trace.err; tail -2 gpurun_out/r2k_e2e_trace.jsonl; grep "trace" gpurun_out/r2k_e2e_trace.err | tail -8
