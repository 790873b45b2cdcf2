set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 --kclique '' > gpurun_out/r2p_bench_4gpu.json 2> gpurun_out/r2p_bench_4gpu.err
grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r2p_bench_4gpu.err | tail -8
cut -c1-300 gpurun_out/r2p_bench_4gpu.json
