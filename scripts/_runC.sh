mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r01j_pytest_gpu.log 2>&1
tail -14 gpurun_out/r01j_pytest_gpu.log
timeout 300 python scripts/prep_bench.py 2>&1 | tee gpurun_out/r01j_prep_bench.txt
