set -x
O=gpurun_out/r2y2; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/p2p_check.py > $O/p2p_check.log 2>&1; echo "p2p rc=$?"; tail -5 $O/p2p_check.log
timeout 900 python -m pytest tests/test_agent_gpu.py tests/test_algos_gpu.py -q -x > $O/agent.log 2>&1; echo "agent rc=$?"; tail -3 $O/agent.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 2 --steps 100 --warmup 20 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r2y2/bench_1gpu.json","gpurun_out/r2y2/bench_2gpu.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value %.1fM ms %.4f e2e %.1fM" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), "opt_us %.1f" % d["breakdown"].get("optimizer_step_kernels_us_per_step"), d["breakdown"].get("optimizer_step_cta0"), d.get("barrier_wait",{}).get("us_per_step_by_rank"), d.get("replicas_identical"), (d.get("peer_equals_nccl") or {}).get("bit_equal"), (d.get("strong") or {}).get("n1_ms_per_step"))
PY
