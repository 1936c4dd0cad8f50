cd $GRAFT_REPO_ROOT
./tools/variants/ldlt_probe
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or tracker" 2>&1 | tail -15
