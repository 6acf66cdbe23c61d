cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or stft_mask_pit or full_size or tie or step or warp_specialised or trainer or review" 2>&1 | tail -3
timeout 600 python tools/variant_bench.py --fused '' --fwd '' --pair 0,1,0,1 2>&1 | grep -v "^$"
