#!/bin/bash
# Round-2 GPU session: parity tests, every bench config, the launch list and one --set full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02_gpu_tests.log 2>&1; tail -5 gpurun_out/r02_gpu_tests.log
for cfg in cfg3 cfg2 cfg4 cfg5-sweep test; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err
  tail -c 600 gpurun_out/r02_bench_$cfg.json; tail -3 gpurun_out/r02_bench_$cfg.err
done
timeout 600 python bench.py --config cfg5-e2e --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg5-e2e.json 2> gpurun_out/r02_bench_cfg5-e2e.err
tail -c 400 gpurun_out/r02_bench_cfg5-e2e.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_persistent -c 1 -f -o gpurun_out/r02_conv3x3_tc_B3360 python tests/probe/ncu_conv_tc.py > gpurun_out/r02_ncu_conv.log 2>&1
tail -3 gpurun_out/r02_ncu_conv.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
tail -2 gpurun_out/r02_launches_bench.log | cut -c1-300
