cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or stft_mask_pit or full_size or tie or step or warp_specialised or trainer or review" 2>&1 | tail -3
for r in 0 1 0 1; do echo "B2S_PAIR_RING=$r"; B2S_PAIR_RING=$r timeout 300 python tools/variant_bench.py --fused '' --fwd '' --pair 1 2>&1 | grep "fused pair"; done
