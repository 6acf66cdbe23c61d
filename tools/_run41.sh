cd $GRAFT_REPO_ROOT
timeout 600 python tools/variant_bench.py --fused '' --fwd '' --ws 0,1,2,0,2 2>&1 | grep -v "^$"
timeout 900 python -m pytest tests -m gpu -x -q -k "warp_specialised" 2>&1 | tail -3
