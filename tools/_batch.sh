cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_dmtet.py -x -q 2>&1 | tail -40 > gpurun_out/b18_dmtet.log
