set -x
O=gpurun_out/r2pdl2; mkdir -p $O
for v in 1 0; do
  RRL_PDL=$v timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_pdl$v.json 2> $O/bench_pdl$v.err; echo "pdl=$v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2pdl2/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1fM ms %.4f e2e %.1fM graph %s pdl %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["config"]["cuda_graph"], d["config"].get("programmatic_dependent_launch")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
timeout 600 python -m pytest tests/test_select_gpu.py -q -k add_both > $O/select.log 2>&1; tail -3 $O/select.log
