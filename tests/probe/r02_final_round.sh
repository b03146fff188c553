#!/bin/bash
# Round-2 final evidence (1 GPU): tests, every bench config (+ reference arm of the default), sanitizer, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02_gpu_tests.log 2>&1; tail -4 gpurun_out/r02_gpu_tests.log
for cfg in cfg3 cfg2 cfg4 cfg5-sweep cfg5-e2e test; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err
  head -c 200 gpurun_out/r02_bench_$cfg.json; echo; tail -2 gpurun_out/r02_bench_$cfg.err | cut -c1-200
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck python tests/probe/sanitize_tc.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tests/probe/sanitize_tc.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_persistent -c 1 -f -o gpurun_out/r02_conv3x3_tc_B3360 python tests/probe/ncu_conv_tc.py > gpurun_out/r02_ncu_conv.log 2>&1; tail -2 gpurun_out/r02_ncu_conv.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tcg_kernel -s 2 -c 1 -f -o gpurun_out/r02_conv_tcg python tests/probe/resnet_bench.py ResNet18 1 > gpurun_out/r02_ncu_tcg.log 2>&1; tail -2 gpurun_out/r02_ncu_tcg.log
