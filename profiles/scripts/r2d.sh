set -x
mkdir -p gpurun_out/r2d
python -m pytest tests/test_agent_gpu.py tests/test_algos_gpu.py tests/test_engine_gpu.py -m gpu -q -x > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
tail -8 gpurun_out/r2d/pytest.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2d/bench_c4.json 2> gpurun_out/r2d/bench_c4.err; echo "bench rc=$?"
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d/launches_warm.csv python profiles/profile_step.py --steps 2 --tc 2 > gpurun_out/r2d/launches_warm.log 2>&1
python profiles/tc_stage_times.py --tc 2 > gpurun_out/r2d/tc_stage_times.txt 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none --warp-sampling-interval 0 -k regex:bwd_tc -c 1 -f -o gpurun_out/r2d/bwd_tc python profiles/profile_step.py --steps 1 --tc 2 > gpurun_out/r2d/ncu_bwd.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none --warp-sampling-interval 0 -k regex:env_step -c 1 -f -o gpurun_out/r2d/env_step python profiles/profile_step.py --steps 1 --tc 2 > gpurun_out/r2d/ncu_env.log 2>&1
head -c 400 gpurun_out/r2d/bench_c4.json
