cd $GRAFT_REPO_ROOT
timeout 400 python tools/dc_geom_probe.py 2>&1 | grep "^RING\|Error" | tail -12
