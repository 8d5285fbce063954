set -x
O=gpurun_out/r2pdl; mkdir -p $O
timeout 900 python -m pytest tests/test_select_gpu.py -q > $O/select.log 2>&1; echo "select rc=$?"; tail -12 $O/select.log
for v in 0 1; do
  RRL_PDL=$v timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_pdl$v.json 2> $O/bench_pdl$v.err; echo "pdl=$v rc=$?"
done
RRL_STAGE_CTAS=116 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_pdl1_s116.json 2> $O/bench_pdl1_s116.err
RRL_STAGE_CTAS=104 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_pdl1_s104.json 2> $O/bench_pdl1_s104.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2pdl/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1fM ms %.4f e2e %.1fM graph %s pdl %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["config"]["cuda_graph"], d["config"].get("programmatic_dependent_launch")))
    except Exception as ex:
        print(f, "ERR", ex)
PY
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_checkpoint_gpu.py tests/test_replay_gpu.py tests/test_env_gpu.py -q -x > $O/engine.log 2>&1; echo "engine rc=$?"; tail -5 $O/engine.log
