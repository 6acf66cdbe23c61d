cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "dc or deep" 2>&1 | tail -2
for i in 1 2; do
B2S_DC_BALANCE=0 timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc " | sed 's/^/balance0 /'
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc " | sed 's/^/balanced /'
done
