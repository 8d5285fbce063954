set -x
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2a/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a/smoke.log
python bench.py --steps 50 --warmup 10 > gpurun_out/r2a/bench_c4.json 2> gpurun_out/r2a/bench_c4.err; echo "bench rc=$?"
python bench.py --steps 50 --warmup 10 --tc 1 --no-cpu-baseline > gpurun_out/r2a/bench_c4_tc1.json 2> gpurun_out/r2a/bench_c4_tc1.err
python bench.py --config C2 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2a/bench_c2.json 2> gpurun_out/r2a/bench_c2.err
python bench.py --config C3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2a/bench_c3.json 2> gpurun_out/r2a/bench_c3.err
python profiles/env_sweep.py --env-name navigation1 > gpurun_out/r2a/env_sweep_nav1.txt 2>&1
python profiles/env_sweep.py --env-name maze --max-log2 20 > gpurun_out/r2a/env_sweep_maze.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a/launches.csv python profiles/profile_step.py --steps 2 --tc 2 > gpurun_out/r2a/launches.log 2>&1
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a/launches_warm.csv python profiles/profile_step.py --steps 2 --tc 2 > gpurun_out/r2a/launches_warm.log 2>&1
python profiles/tc_stage_times.py --tc 2 > gpurun_out/r2a/tc_stage_times.txt 2>&1
head -c 1500 gpurun_out/r2a/bench_c4.json
