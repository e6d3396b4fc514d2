cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dmtet.py -x -q -s -k "reference_run_dmtet" 2>&1 | grep -v "^$" | tail -45 > gpurun_out/b29_dmtet.log
