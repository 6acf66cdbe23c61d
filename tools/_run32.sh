cd $GRAFT_REPO_ROOT
timeout 600 python tools/variant_bench.py --fused '' --fwd '' --pair 0,1,0,1 2>&1 | grep -v "^$"
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or stft_mask_pit or full_size or tie or step" 2>&1 | tail -5
