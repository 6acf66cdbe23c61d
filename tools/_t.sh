cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "dc or deep" 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc "
timeout 300 python bench.py --config dc 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dc config', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_probe2.py > gpurun_out/san4_racecheck.log 2>&1; tail -2 gpurun_out/san4_racecheck.log; grep -o "and Read access at void <unnamed>::[a-z_0-9]*\|Write access at [a-z:_0-9A-Z]*" gpurun_out/san4_racecheck.log | sort | uniq -c
