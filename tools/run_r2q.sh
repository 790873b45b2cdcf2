set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_build.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -4 gpurun_out/r2q_pytest.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --kclique '' > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err
grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r2q_bench_2gpu.err | tail -6
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_2gpu.json')); print(d['ms_per_step'], d['schedule_ms'], d['count_ms'], d['e2e'])"
