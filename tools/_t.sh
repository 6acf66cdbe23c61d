cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "dc or deep" 2>&1 | tail -2
B2S_DC_RING=0 timeout 300 python tools/dc_geom_probe.py 2>&1 | grep "^RING"
B2S_DC_RING=1 timeout 300 python tools/dc_geom_probe.py 2>&1 | grep "^RING"
timeout 300 python tools/kernel_bench.py 2>&1 | grep "^dc "
