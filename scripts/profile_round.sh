#!/bin/bash
# Round profile: (1) launch list of the bench command, (2) ncu --set full of the dominant kernel in the
# bench configuration.  Run on the GPU box: gpurun -- 'bash scripts/profile_round.sh r01g'
tag=${1:-r01x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:icp_persistent --launch-skip 3 --launch-count 1 \
    -o gpurun_out/${tag}_icp_bench -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench_line.json
