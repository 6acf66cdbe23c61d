cd $GRAFT_REPO_ROOT
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitizer_probe.py > gpurun_out/san_mem2.log 2>&1; tail -4 gpurun_out/san_mem2.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitizer_probe.py > gpurun_out/san_race2.log 2>&1; tail -4 gpurun_out/san_race2.log
