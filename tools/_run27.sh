cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "targets or warp_specialised" 2>&1 | tail -5
