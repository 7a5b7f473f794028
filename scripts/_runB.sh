mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_pair or reference_kdtree" 2>&1 | tail -3
timeout 600 python scripts/ab_lean.py piecewise-icp_b200/libpwicp.so libpwicp_t128.so libpwicp_t384.so libpwicp_t512.so libpwicp_t768.so libpwicp_t256x2.so piecewise-icp_b200/libpwicp.so 2>&1 | tee gpurun_out/r01i_ab_cta.txt
