cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_sd_gemm.py -q -x -k "stride2 or conv3x3" > gpurun_out/c8_conv_tests.log 2>&1
tail -6 gpurun_out/c8_conv_tests.log
timeout 400 python -m pytest tests/test_gpu_sd_engine.py tests/test_gpu_zero123.py -q -x > gpurun_out/c8_engine_tests.log 2>&1
tail -4 gpurun_out/c8_engine_tests.log
timeout 600 python tools/bench_lists.py > gpurun_out/c8_lists.log 2>&1
tail -4 gpurun_out/c8_lists.log
