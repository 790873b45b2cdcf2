set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_full_scale.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log; tail -3 gpurun_out/r2t_pytest.log
timeout 100 python tools/partition_balance.py --scale 24 --parts 1,8 --reps 3 > gpurun_out/r2t_partition_balance.jsonl 2>&1; cut -c1-400 gpurun_out/r2t_partition_balance.jsonl
