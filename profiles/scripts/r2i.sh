set -x
mkdir -p gpurun_out/r2i
python -m pytest tests/test_agent_gpu.py tests/test_engine_gpu.py -m gpu -q -x > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/pytest.log
tail -6 gpurun_out/r2i/pytest.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2i/bench_c4.json 2> gpurun_out/r2i/bench_c4.err; echo "bench rc=$?"
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i/launches_warm.csv python profiles/profile_step.py --steps 2 --tc 2 > gpurun_out/r2i/launches_warm.log 2>&1
head -c 300 gpurun_out/r2i/bench_c4.json
