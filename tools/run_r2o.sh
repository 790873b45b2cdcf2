set -x
mkdir -p gpurun_out
timeout 300 python tools/partition_balance.py --scale 24 --parts 1,2,4,8 > gpurun_out/r2o_partition_balance.jsonl 2> gpurun_out/r2o_partition_balance.err
timeout 300 python tools/partition_balance.py --scale 24 --parts 1,8 --row-walk >> gpurun_out/r2o_partition_balance.jsonl 2>> gpurun_out/r2o_partition_balance.err
cut -c1-330 gpurun_out/r2o_partition_balance.jsonl; tail -3 gpurun_out/r2o_partition_balance.err
GMSB_TC_TRACE=1 timeout 200 python tools/tc_sweep.py --scale 24 --reps 2 --configs '[{}]' > gpurun_out/r2o_trace.jsonl 2> gpurun_out/r2o_trace.err; grep trace gpurun_out/r2o_trace.err | tail -16
