cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dmtet.py -x -q -s -k "antialias or training" 2>&1 | tail -40 > gpurun_out/b20_dmtet.log
