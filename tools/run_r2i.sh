set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2i_bench_8gpu.json 2> gpurun_out/r2i_bench_8gpu.err
grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r2i_bench_8gpu.err | tail -12
cut -c1-400 gpurun_out/r2i_bench_8gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --scale 26 --rmat-a 0.65 --kclique '' --no-e2e > gpurun_out/r2i_bench_s26_8gpu.json 2> gpurun_out/r2i_bench_s26_8gpu.err
grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r2i_bench_s26_8gpu.err | tail -6
cut -c1-600 gpurun_out/r2i_bench_s26_8gpu.json
GMSB_DEVICES=0,1,2,3,4,5,6,7 timeout 300 oracle/_ref/dropin_tc -g kronecker 20 --deg 16 -n 2 -v > gpurun_out/r2i_dropin_8gpu.log 2>&1; grep "@@@\|devices" gpurun_out/r2i_dropin_8gpu.log | grep -v SortedSetGraph
