#!/bin/bash
# Round profile: (1) launch list of the bench command, (2) ncu --set full of the dominant kernel in the
# bench configuration, (3) the same at 10M (BASELINE configs[4]), (4) the bench line.
# Run on the GPU box: gpurun -- 'bash scripts/profile_round.sh r02w'
tag=${1:-r02x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
# the persistent kernel runs as two launches per step (iteration 0 | iterations 1..49): both of the second step, plus the
# two search kernels around them (icp_seed_kernel = iteration-0 pre-pass, icp_research_kernel = iteration 1)
ncu --set full --import-source on --clock-control none -k regex:"icp_persistent|icp_research|icp_seed" --launch-skip 4 --launch-count 4 \
    -o gpurun_out/${tag}_icp_bench -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:icp_persistent --launch-skip 5 --launch-count 1 \
    -o gpurun_out/${tag}_icp_10m -f python scripts/prof_icp_only.py 10000000 50 > gpurun_out/${tag}_ncu_10m.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
tail -c 800 gpurun_out/${tag}_bench_line.json
