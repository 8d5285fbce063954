set -x
mkdir -p gpurun_out/r2b
python -m pytest tests -m gpu -q > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -15 gpurun_out/r2b/pytest.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2b/bench_c4.json 2> gpurun_out/r2b/bench_c4.err; echo "bench rc=$?"
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b/launches_warm.csv python profiles/profile_step.py --steps 2 --tc 2 > gpurun_out/r2b/launches_warm.log 2>&1
python profiles/tc_stage_times.py --tc 2 > gpurun_out/r2b/tc_stage_times.txt 2>&1
head -c 600 gpurun_out/r2b/bench_c4.json
