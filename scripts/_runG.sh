mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01k_pytest_gpu.log 2>&1; tail -6 gpurun_out/r01k_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r01k_bench_line.json 2> gpurun_out/r01k_bench.err; tail -c 900 gpurun_out/r01k_bench_line.json
