set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_sharded_build.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --kclique '' > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err
grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r2l_bench_2gpu.err | tail -12
cut -c1-300 gpurun_out/r2l_bench_2gpu.json
GMSB_DEVICES=0,1 timeout 200 oracle/_ref/dropin_tc -g kronecker 18 --deg 16 -n 2 -v > gpurun_out/r2l_dropin_2gpu.log 2>&1; grep "@@@\|devices" gpurun_out/r2l_dropin_2gpu.log | grep -v SortedSetGraph
