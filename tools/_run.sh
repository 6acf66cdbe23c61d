cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "dc or deep" 2>&1 | tail -2
timeout 400 python tools/dc_geom_probe.py 2>&1 | grep "^RING"
