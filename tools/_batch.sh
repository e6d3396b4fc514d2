cd $GRAFT_REPO_ROOT
timeout 95 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02_session2.csv python tools/profile_step.py lambertian > gpurun_out/c15_profile.log 2>&1
tail -2 gpurun_out/c15_profile.log
