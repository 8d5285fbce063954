set -x
mkdir -p gpurun_out/r2j
python -m pytest tests/test_engine_gpu.py tests/test_checkpoint_gpu.py tests/test_dropin_gpu.py -m gpu -q > gpurun_out/r2j/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest.log
tail -6 gpurun_out/r2j/pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2j/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2j/bench_c4.json 2> gpurun_out/r2j/bench_c4.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j/bench_c4_driver.json 2> gpurun_out/r2j/bench_c4_driver.err; echo "bench rc=$?"
head -c 300 gpurun_out/r2j/bench_c4.json
