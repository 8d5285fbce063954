# Round-2 evidence run (one B200): bench lines of the BASELINE configs, env N-sweeps, launch lists, ncu --set full captures.
set -x
O=gpurun_out/r2p
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log; tail -4 $O/pytest_full.log
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
python bench.py --steps 200 --warmup 20 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --config C2 --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --config C3 --steps 200 --warmup 20 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --config C5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
python profiles/env_sweep.py --env-name navigation1 > $O/env_sweep_nav1.txt 2>&1
python profiles/env_sweep.py --env-name navigation2 > $O/env_sweep_nav2.txt 2>&1
python profiles/env_sweep.py --env-name maze --max-log2 20 > $O/env_sweep_maze.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python profiles/profile_step.py --steps 2 --tc 2 > $O/launches.log 2>&1
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_warm.csv python profiles/profile_step.py --steps 2 --tc 2 > $O/launches_warm.log 2>&1
python profiles/tc_stage_times.py --tc 2 > $O/tc_stage_times.txt 2>&1
for k in act_tc fwd_tc bwd_tc adam_tile env_step replay_sample; do
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o $O/ncu_$k python profiles/profile_step.py --steps 1 --tc 2 > $O/ncu_$k.log 2>&1
done
ls -la $O
