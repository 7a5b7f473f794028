mkdir -p gpurun_out
timeout 300 python scripts/prep_bench.py 2>&1 | tee gpurun_out/r01j_prep_bench.txt
cat > /tmp/knn1m.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "piecewise-icp_b200", "python"))
import pwicp_b200 as P
from pwicp_b200 import synth
c = synth.make_pair(1000000, with_clouds=False)["ct1"]
ctx = P.Context(0)
for _ in range(3):
    ctx.knn_mean_dist(c, 14); print(ctx.last_device_ms(), ctx.last_knn_kernel_ms())
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_mean_dist --launch-skip 2 --launch-count 1 -o gpurun_out/r01j_knn -f python /tmp/knn1m.py > gpurun_out/r01j_knn_ncu.log 2>&1
tail -5 gpurun_out/r01j_knn_ncu.log
