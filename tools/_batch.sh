# scratch batch for `gpurun -- 'bash tools/_batch.sh'`: the end-of-round verification (tests, bench line, launch list)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; tail -3 gpurun_out/final_gpu_tests.log
timeout 600 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; tail -c 400 gpurun_out/final_bench_c2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py lambertian > gpurun_out/final_profile.log 2>&1; tail -1 gpurun_out/final_profile.log
