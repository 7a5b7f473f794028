# A/B of several builds of libpwicp.so on one box: bash scripts/run_ab.sh <tag> [n] [iters] -- variants are ../libpwicp_<name>.so
tag=${1:-ab}; n=${2:-1000000}; iters=${3:-50}
mkdir -p gpurun_out; : > gpurun_out/${tag}.txt
for lib in "" $(ls libpwicp_*.so 2>/dev/null); do
  echo "=== variant ${lib:-main}" >> gpurun_out/${tag}.txt
  PWICP_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/icp_probe.py $n $iters >> gpurun_out/${tag}.txt 2>&1
done
grep -v "^parity" gpurun_out/${tag}.txt
