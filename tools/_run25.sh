cd $GRAFT_REPO_ROOT
for r in 1 2; do for l in a b; do cp ab/lib_$l.so padertorch_b200/libb200sep.so; echo "lib_$l"; timeout 300 python tools/variant_bench.py --fused 0 --fwd '' 2>&1 | grep -v "^$"; done; done
cp ab/lib_b.so padertorch_b200/libb200sep.so
for s in 148 111 74 37; do echo "SMS=$s"; B2S_FUSED_SMS=$s timeout 300 python tools/variant_bench.py --fused 0 --fwd '' 2>&1 | grep -v "^$"; done
