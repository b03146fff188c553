#!/bin/bash
# Evidence for the kernels added last: sanitizer, the ResNet50 end-to-end point, the cfg4 launch list.
mkdir -p gpurun_out
timeout -k 5 150 compute-sanitizer --tool memcheck python tests/probe/sanitize_tc.py stem s2 > gpurun_out/r02_sanitizer_memcheck_stem_s2.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck_stem_s2.log
timeout -k 5 150 compute-sanitizer --tool racecheck python tests/probe/sanitize_tc.py stem s2 > gpurun_out/r02_sanitizer_racecheck_stem_s2.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck_stem_s2.log
timeout -k 5 120 python bench.py --config cfg5-e2e --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg5-e2e.json 2> gpurun_out/r02_bench_cfg5-e2e.err; head -c 150 gpurun_out/r02_bench_cfg5-e2e.json; echo
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_cfg4.log 2>&1
wc -l gpurun_out/r02_launches_cfg4.csv
