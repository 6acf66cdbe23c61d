cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/pair_probe.py 2>&1 | grep "^B="
B2S_PAIR_CLUSTER=0 timeout 200 python tools/pair_probe.py 2>&1 | grep "^B=64"
