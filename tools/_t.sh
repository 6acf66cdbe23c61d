cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "stft or fused or full_size or golden" 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^stft |Y|\|^stft complex\|fused STFT->PIT (reads\|istft backward"
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^stft |Y|\|^stft complex\|fused STFT->PIT (reads\|istft backward"
