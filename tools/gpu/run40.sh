cd $GRAFT_REPO_ROOT
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_sfm_run_step_matches_oracle and (160-120-32 or 80-60-32)" 2>&1 | tail -4
