cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "dc or deep_clustering or clustering" 2>&1 | tail -3
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc "
