set -x
mkdir -p gpurun_out/r2s
timeout 1500 python -m pytest tests/test_select_gpu.py -q -x > gpurun_out/r2s/select.log 2>&1; echo "select rc=$?"; tail -30 gpurun_out/r2s/select.log
timeout 1200 python -m pytest tests/test_dropin_gpu.py -q -k "vectorised" > gpurun_out/r2s/dropin_vec.log 2>&1; echo "dropin rc=$?"; tail -8 gpurun_out/r2s/dropin_vec.log
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_algos_gpu.py -q > gpurun_out/r2s/engine.log 2>&1; echo "engine rc=$?"; tail -5 gpurun_out/r2s/engine.log
