#!/bin/bash
# One GPU-box visit: feeder parity tests, feeder bench line, ncu --set full capture of the feeder kernel, default bench.
mkdir -p gpurun_out
python -m pytest tests/test_feeder_gpu.py -x -q > gpurun_out/feeder_tests.log 2>&1; tail -5 gpurun_out/feeder_tests.log
python bench.py --workload feeder --steps 10 > gpurun_out/feeder_bench.json 2> gpurun_out/feeder_bench.err
cat gpurun_out/feeder_bench.json; tail -3 gpurun_out/feeder_bench.err
ncu --set full --clock-control none --import-source on -k regex:episode_transform --launch-skip 5 -c 1 \
    -o gpurun_out/r01_episode_transform -f python bench.py --workload feeder --steps 3 --no-cpu-baseline > gpurun_out/ncu_feed.log 2>&1
tail -2 gpurun_out/ncu_feed.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
cat gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
