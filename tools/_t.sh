cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc \|tasnet\|^pit_sse forward"
timeout 200 python tools/pair_probe.py 2>&1 | grep "^B=64"
