cd $GRAFT_REPO_ROOT
export DFK_LIB=$GRAFT_REPO_ROOT/tools/variants/libdfk_wd.so
timeout 60 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 2 --sustain-seconds 0.1 > gpurun_out/r2_t18.out 2>gpurun_out/r2_t18.err
echo rc $?
sort gpurun_out/r2_t18.out | uniq -c | sort -rn | head -30
tail -5 gpurun_out/r2_t18.err
