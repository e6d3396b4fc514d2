cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dmtet.py -x -q 2>&1 | tail -25 > gpurun_out/b23_dmtet.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02_c5.csv python tools/profile_step.py lambertian dmtet > gpurun_out/b23_profile.log 2>&1
tail -3 gpurun_out/b23_profile.log
