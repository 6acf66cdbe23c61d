cd $GRAFT_REPO_ROOT
for r in 1 2; do for l in a b c d; do cp ab/lib_$l.so padertorch_b200/libb200sep.so; echo "lib_$l"; timeout 300 python tools/variant_bench.py --fused '' --fwd '' --pair 1 2>&1 | grep "fused pair"; done; done
