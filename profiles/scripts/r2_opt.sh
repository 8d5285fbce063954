set -x
mkdir -p gpurun_out/r2o
timeout 600 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/r2o/bench_1gpu.json 2> gpurun_out/r2o/bench_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 2 --steps 100 --warmup 20 --no-checks --no-cpu-baseline > gpurun_out/r2o/bench_2gpu.json 2> gpurun_out/r2o/bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r2o/bench_1gpu.json","gpurun_out/r2o/bench_2gpu.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.4f" % d["ms_per_step"], "opt_us", d["breakdown"].get("optimizer_step_kernels_us_per_step"), d["breakdown"].get("optimizer_step_cta0"), d.get("barrier_wait"))
PY
