cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "tasnet or pair or loss_set or losses" 2>&1 | tail -2
B2S_PAIR_THREADS=128 timeout 600 python -m pytest tests -m gpu -x -q -k "tasnet or pair or loss_set or losses" 2>&1 | tail -2
timeout 200 python tools/pair_probe.py 2>&1 | grep "^B="
B2S_PAIR_THREADS=128 timeout 200 python tools/pair_probe.py 2>&1 | grep "^B="
B2S_PAIR_THREADS=128 B2S_PAIR_CTAS=3 timeout 200 python tools/pair_probe.py 2>&1 | grep "^B="
