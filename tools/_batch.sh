cd $GRAFT_REPO_ROOT
timeout 600 python tools/bench_lists.py > gpurun_out/c10_lists.log 2>&1
tail -3 gpurun_out/c10_lists.log
