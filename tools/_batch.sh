cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_gpu_trainer.py -x -q > gpurun_out/c13_trainer.log 2>&1; tail -3 gpurun_out/c13_trainer.log
timeout 150 python -m pytest tests/test_gpu_dropin_reference.py -x -q > gpurun_out/c13_dropin.log 2>&1; tail -3 gpurun_out/c13_dropin.log
timeout 100 python -m pytest tests/test_gpu_dmtet.py -x -q > gpurun_out/c13_dmtet.log 2>&1; tail -3 gpurun_out/c13_dmtet.log
