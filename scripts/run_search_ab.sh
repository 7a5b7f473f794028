mkdir -p gpurun_out; : > gpurun_out/r02j_search.txt
for lib in "" libpwicp_legacy.so libpwicp_all.so; do
  echo "=== variant ${lib:-main}" >> gpurun_out/r02j_search.txt
  PWICP_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/search_probe.py 1000000 >> gpurun_out/r02j_search.txt 2>&1
  PWICP_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/icp_probe.py 1000000 50 2>&1 | grep -v "^parity n=1980" >> gpurun_out/r02j_search.txt
done
cat gpurun_out/r02j_search.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
