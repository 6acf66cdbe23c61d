cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_properties" 2>&1 | grep -v "^$" | tail -40
