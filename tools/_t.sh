cd $GRAFT_REPO_ROOT
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_probe2.py > gpurun_out/san2_memcheck.log 2>&1; tail -4 gpurun_out/san2_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitizer_probe2.py > gpurun_out/san2_racecheck.log 2>&1; tail -4 gpurun_out/san2_racecheck.log
