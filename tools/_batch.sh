cd $GRAFT_REPO_ROOT
timeout 900 python tools/bench_lists.py > gpurun_out/c4_lists.log 2>&1
cat gpurun_out/c4_lists.log | tail -8
