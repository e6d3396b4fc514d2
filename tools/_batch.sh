cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_field_marched python tools/profile_kernels.py marched field-only > gpurun_out/b33_ncu.log 2>&1
tail -3 gpurun_out/b33_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02_final.csv python tools/profile_step.py lambertian > gpurun_out/b33_profile.log 2>&1
tail -2 gpurun_out/b33_profile.log
