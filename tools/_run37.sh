cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for p in 0 1 0 1; do echo "B2S_FWD_PAIR=$p"; B2S_FWD_PAIR=$p timeout 300 python tools/kernel_bench.py 2>&1 | grep "^stft \|istft backward\|prepare"; done
