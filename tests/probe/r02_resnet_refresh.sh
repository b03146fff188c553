#!/bin/bash
# Round-2 refresh after the ResNet work (1 GPU): tests, ResNet bench configs, sanitizer on the new kernels, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02_gpu_tests.log 2>&1; tail -4 gpurun_out/r02_gpu_tests.log
for cfg in cfg4 cfg5-e2e; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err
  head -c 200 gpurun_out/r02_bench_$cfg.json; echo; tail -2 gpurun_out/r02_bench_$cfg.err | cut -c1-200
done
timeout 900 compute-sanitizer --tool memcheck python tests/probe/sanitize_tc.py tcg stem > gpurun_out/r02_sanitizer_memcheck_resnet.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck_resnet.log
timeout 900 compute-sanitizer --tool racecheck python tests/probe/sanitize_tc.py tcg stem > gpurun_out/r02_sanitizer_racecheck_resnet.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck_resnet.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tcg_kernel|stem_tc_kernel" -c 3 -f -o gpurun_out/r02_conv_tcg python tests/probe/ncu_conv_tcg.py > gpurun_out/r02_ncu_tcg.log 2>&1; tail -2 gpurun_out/r02_ncu_tcg.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_cfg4.log 2>&1
python tests/probe/time_conv_tcg.py > gpurun_out/r02_conv_tcg_layers.log 2>&1
