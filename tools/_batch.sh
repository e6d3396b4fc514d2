cd $GRAFT_REPO_ROOT
timeout 40 python -m pytest tests/test_gpu_render_aux.py tests/test_gpu_adan.py -x -q > gpurun_out/c16_aux.log 2>&1; tail -3 gpurun_out/c16_aux.log
