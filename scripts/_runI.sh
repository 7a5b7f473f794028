mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r01l_pytest_gpu.log 2>&1; tail -5 gpurun_out/r01l_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/steady_10m.py 2>&1 | tee gpurun_out/r01l_10m.txt | tail -1
