cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "dc or deep_clustering or clustering" 2>&1 | tail -3
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc "
timeout 300 python bench.py --config dc --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
