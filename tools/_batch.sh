set -x
cd $GRAFT_REPO_ROOT
B="--steps 40 --warmup 5 --no-ref-cuda --no-eval --no-cpu-baseline"
(cd _old && timeout 200 python bench.py $B > ../gpurun_out/b16_bench_old.log 2>&1)
timeout 200 python bench.py $B > gpurun_out/b16_bench_new.log 2>&1
SDF_DIRECT_CONV_IN=0 timeout 200 python bench.py $B > gpurun_out/b16_bench_new_noconv.log 2>&1
(cd _old && timeout 200 python bench.py $B > ../gpurun_out/b16_bench_old2.log 2>&1)
timeout 100 python tools/bench_conv_in.py > gpurun_out/b16_conv.log 2>&1
timeout 300 python tools/profile_ops.py 10 > gpurun_out/b16_ops_new.txt 2>&1
(cd _old && timeout 300 python tools/profile_ops.py 10 > ../gpurun_out/b16_ops_old.txt 2>&1)
timeout 300 python -m pytest tests/test_gpu_sd_ops.py -x -q 2>&1 | tail -5 > gpurun_out/b16_tests.log
grep -h -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus"' gpurun_out/b16_bench_*.log
