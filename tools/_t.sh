cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "dc or deep" 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc "
timeout 300 python bench.py --config dc 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dc config', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('step_roofline'))"
