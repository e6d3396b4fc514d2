set -x
cd $GRAFT_REPO_ROOT
B="--steps 40 --warmup 5 --no-ref-cuda --no-eval --no-cpu-baseline"
timeout 200 python bench.py $B > gpurun_out/b17_bench_new.log 2>&1
(cd _old && timeout 200 python bench.py $B > ../gpurun_out/b17_bench_old.log 2>&1)
SDF_GN_CLUSTER=1 timeout 200 python bench.py $B > gpurun_out/b17_bench_cluster.log 2>&1
timeout 200 python bench.py $B > gpurun_out/b17_bench_new2.log 2>&1
SDF_GN_CLUSTER=1 timeout 300 python tools/profile_ops.py 10 2>&1 | head -14 > gpurun_out/b17_ops_cluster.txt
SDF_GN_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_sd_ops.py -x -q 2>&1 | tail -5 > gpurun_out/b17_tests_cluster.log
timeout 600 python -m pytest tests/test_gpu_sd_ops.py tests/test_gpu_sd_gemm.py tests/test_gpu_sd_engine.py -x -q 2>&1 | tail -5 > gpurun_out/b17_tests.log
grep -h -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus"' gpurun_out/b17_bench_*.log
