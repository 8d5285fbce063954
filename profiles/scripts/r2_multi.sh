# multi-GPU bench line exactly as the driver launches it: bash profiles/scripts/r2_multi.sh N
set -x
N=$1
mkdir -p gpurun_out/r2m
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/r2m/bench_${N}gpu.json 2> gpurun_out/r2m/bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r2m/bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29771 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2m/bench_${N}gpu_reference.json 2> gpurun_out/r2m/bench_${N}gpu_reference.err; echo "ref rc=$?"
head -c 400 gpurun_out/r2m/bench_${N}gpu.json
