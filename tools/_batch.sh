cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/b31_gpu_tests.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/b31_bench.log 2>&1
timeout 300 python bench.py --config C5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b31_bench_c5.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/b31_smoke.log 2>&1
tail -3 gpurun_out/b31_gpu_tests.log; tail -2 gpurun_out/b31_smoke.log
