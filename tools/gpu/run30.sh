cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sfm_step_tc -s 4 -c 1 -f -o gpurun_out/prof_r02_v5_sfm_step_tc python bench.py --steps 2 --warmup 1 --no-verify --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0 > /dev/null 2> gpurun_out/r2_t30_ncu.err
ls -la gpurun_out/prof_r02_v5_sfm_step_tc.ncu-rep
