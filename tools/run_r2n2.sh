set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_sharded_build.py tests/test_cpp_facade.py -m gpu -x -q > gpurun_out/r2n_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_2gpu.log; tail -4 gpurun_out/r2n_pytest_2gpu.log
