#!/bin/bash
# Round-2 multi-GPU evidence on one 8 x B200 box: NCCL test, cfg3 at 8, cfg4 and the cfg5 sweep at 2 / 4 / 8 ranks.
mkdir -p gpurun_out
./tests/probe/bin/umma_probe ts_bf16 > gpurun_out/r02_probe_ts_bf16.log 2>&1; tail -2 gpurun_out/r02_probe_ts_bf16.log
timeout 600 python -m pytest tests/test_multi_nccl_gpu.py -m gpu -q 2>&1 | tail -3
run() {  # n config steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
    bench.py --gpus $1 --config $2 --steps $3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02_scale_$2_n$1.err | tail -1 > gpurun_out/r02_scale_$2_n$1.json
  head -c 260 gpurun_out/r02_scale_$2_n$1.json; echo; tail -2 gpurun_out/r02_scale_$2_n$1.err | cut -c1-200
}
run 8 cfg3 10
for n in 2 4 8; do run $n cfg4 5; run $n cfg5-sweep 10; done
