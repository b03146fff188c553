#!/bin/bash
# Last GPU call of round 2: the stem weight-gradient kernel alone first (hard time limit), then the whole GPU suite and the
# bench lines of the two headline configurations.  Logs go straight to files (nothing is lost if a step is killed).
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests -m gpu -x -q -k "stem_tc" > gpurun_out/f_stem.log 2>&1
echo "stem rc=$?" > gpurun_out/f_stem.rc
tail -2 gpurun_out/f_stem.log
grep -q "4 passed" gpurun_out/f_stem.log || { echo "stem kernels failed: not running the rest"; exit 1; }
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
timeout -k 5 240 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; head -c 150 gpurun_out/r02_bench_cfg3.json; echo
timeout -k 5 240 python bench.py --config cfg4 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; head -c 150 gpurun_out/r02_bench_cfg4.json; echo
cat gpurun_out/f_stem.rc
