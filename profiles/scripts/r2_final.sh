# Final round-2 evidence (one B200), code state = HEAD: bench lines, launch lists, ncu of the acting kernel.
set -x
O=gpurun_out/r2z
mkdir -p $O
python bench.py --steps 200 --warmup 20 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --config C2 --steps 200 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --config C3 --steps 200 --warmup 20 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --config C5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 100 --warmup 20 --tc 1 --no-cpu-baseline > $O/bench_c4_tc1.json 2> $O/bench_c4_tc1.err
python bench.py --steps 50 --warmup 10 --tc 0 --no-cpu-baseline > $O/bench_c4_tc0.json 2> $O/bench_c4_tc0.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python profiles/profile_step.py --steps 2 --tc 2 > $O/launches.log 2>&1
ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_warm.csv python profiles/profile_step.py --steps 2 --tc 2 > $O/launches_warm.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:act_tc -c 3 -f -o $O/ncu_act_tc python profiles/profile_step.py --steps 1 --tc 2 > $O/ncu_act_tc.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:act_tc -c 1 -f -o $O/ncu_act_tc_all python profiles/time_kernels.py --steps 6 > $O/ncu_act_tc_all.log 2>&1
ls -la $O
