set -x
mkdir -p gpurun_out/r2y
for m in 0 1 2 3; do
RRL_SYNC_MODE=$m timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2976$m bench.py --gpus 2 --steps 100 --warmup 20 --no-checks --no-cpu-baseline > gpurun_out/r2y/bench_2gpu_m$m.json 2> gpurun_out/r2y/bench_2gpu_m$m.err
done
python - <<'PY'
import json
for m in range(4):
    f="gpurun_out/r2y/bench_2gpu_m%d.json"%m
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(m, "ms %.4f" % d["ms_per_step"], "opt_us %.1f" % d["breakdown"].get("optimizer_step_kernels_us_per_step"), d["breakdown"].get("optimizer_step_cta0"), d.get("barrier_wait",{}).get("us_per_step_by_rank"))
    except Exception as ex: print(m, "ERR", ex)
PY
