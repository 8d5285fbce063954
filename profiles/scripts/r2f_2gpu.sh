set -x
mkdir -p gpurun_out/r2f
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/p2p_check.py > gpurun_out/r2f/p2p_check.log 2>&1; echo "p2p rc=$?"; tail -6 gpurun_out/r2f/p2p_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/r2f/bench_2gpu.json 2> gpurun_out/r2f/bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2f/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 50 --warmup 10 --peer-grads 0 --no-checks > gpurun_out/r2f/bench_2gpu_nccl.json 2> gpurun_out/r2f/bench_2gpu_nccl.err; echo "bench2 nccl rc=$?"
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2f/bench_1gpu.json 2> gpurun_out/r2f/bench_1gpu.err
head -c 3000 gpurun_out/r2f/bench_2gpu.json
