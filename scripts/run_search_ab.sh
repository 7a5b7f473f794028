# A/B of the search-side probes across builds: bash scripts/run_search_ab.sh [lib ...]  ("" = the in-tree build)
for lib in "" "$@"; do
  echo "=== variant ${lib:-main}"
  PWICP_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/search_probe.py 300000 2>&1 | tail -7
  PWICP_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/icp_probe.py 1000000 50 2>&1 | grep "^timing\|iteration us"
done
