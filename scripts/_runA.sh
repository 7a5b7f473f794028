mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r01i_pytest_gpu.log 2>&1
tail -12 gpurun_out/r01i_pytest_gpu.log
( time timeout 600 python scripts/refdata_4d.py run --refseg --modes 0 --epochs 8 ) > gpurun_out/r01i_refdata_refseg.log 2>&1
tail -4 gpurun_out/r01i_refdata_refseg.log
cat gpurun_out/refdata_4d_refseg_report.txt
