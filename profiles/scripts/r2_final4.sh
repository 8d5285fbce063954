set -x
O=gpurun_out/r2f4; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log; tail -4 $O/pytest_full.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f4/bench_c4.json").read().strip().splitlines()[-1])
print("value %.1fM ms %.4f e2e %.1fM roofline %.4f cpu %s launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["roofline"]["frac"], d["cpu_baseline"]["value"], d["gpu_launches"]), d["breakdown"])
PY
