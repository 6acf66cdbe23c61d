cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc \|^pit_sse"
B2S_DC_BWD_STAGES=2 timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc backward" | sed 's/^/two stages: /'
timeout 300 python bench.py --config dc 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dc config', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['step_roofline']['frac'])"
timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_probe2.py > gpurun_out/san5_racecheck.log 2>&1; tail -1 gpurun_out/san5_racecheck.log
