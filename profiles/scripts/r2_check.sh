set -x
mkdir -p gpurun_out/r2c2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/p2p_check.py > gpurun_out/r2c2/p2p_check.log 2>&1; echo "p2p rc=$?"; tail -8 gpurun_out/r2c2/p2p_check.log
python -m pytest tests -m gpu -q > gpurun_out/r2c2/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2/pytest_full.log; tail -5 gpurun_out/r2c2/pytest_full.log
