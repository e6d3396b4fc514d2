cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --config C5 --steps 20 --warmup 4 --no-cpu-baseline --breakdown > gpurun_out/b22_bench_c5.log 2>&1
tail -c 3000 gpurun_out/b22_bench_c5.log
