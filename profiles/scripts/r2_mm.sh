set -x
N=$1
O=gpurun_out/r2mm; mkdir -p $O
if [ "$N" = "2" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/p2p_check.py > $O/p2p_check.log 2>&1; echo "p2p rc=$?"; tail -5 $O/p2p_check.log
fi
for mm in 1 0; do
RRL_MULTIMEM=$mm timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2976$mm bench.py --gpus $N --steps 100 --warmup 20 --no-cpu-baseline > $O/bench_${N}gpu_mm$mm.json 2> $O/bench_${N}gpu_mm$mm.err; echo "rc=$?"; tail -2 $O/bench_${N}gpu_mm$mm.err
done
python - <<PY
import json
for mm in (1,0):
    f="gpurun_out/r2mm/bench_${N}gpu_mm%d.json"%mm
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(mm, "value %.1fM ms %.4f e2e %.1fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), d["config"]["grad_allreduce"][:40], "opt_us %.1f" % d["breakdown"].get("optimizer_step_kernels_us_per_step"), d["breakdown"].get("optimizer_step_cta0"), d.get("barrier_wait",{}).get("us_per_step_by_rank"), d.get("replicas_identical"), d.get("peer_equals_nccl"), (d.get("strong") or {}).get("n1_ms_per_step"))
    except Exception as ex: print(mm, "ERR", ex)
PY
