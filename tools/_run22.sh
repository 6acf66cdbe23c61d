cd $GRAFT_REPO_ROOT
timeout 600 python tools/variant_bench.py --fused 0 --fwd 0 --ablate 0,1,4,5,7,13,15,2 2>&1 | grep -v "^$"
