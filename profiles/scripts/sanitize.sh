# compute-sanitizer pass: memcheck + racecheck over small cases of every kernel family, through the C ABI.
# Output: gpurun_out/sanitize/*.log (copied to profiles/sanitizer_r2_*.txt)
set -x
mkdir -p gpurun_out/sanitize
SAN=/usr/local/cuda/bin/compute-sanitizer
T="timeout 900"
AGENT='tests/test_agent_gpu.py::test_updates_and_acting_vs_reference[maze_b64-2] tests/test_agent_gpu.py::test_updates_and_acting_vs_reference[maze_b64-0]'
OTHERS='tests/test_env_gpu.py tests/test_replay_gpu.py::test_sample_streams_bit_exact_vs_reference tests/test_replay_gpu.py::test_gates_and_stream_not_advanced_when_closed'
for tool in memcheck racecheck; do
  $T $SAN --tool $tool --print-limit 50 --error-exitcode 0 python -m pytest $AGENT -m gpu -q -x > gpurun_out/sanitize/${tool}_agent.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize/${tool}_agent.log
  $T $SAN --tool $tool --print-limit 50 --error-exitcode 0 python -m pytest $OTHERS -m gpu -q -x > gpurun_out/sanitize/${tool}_env_replay.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize/${tool}_env_replay.log
  tail -4 gpurun_out/sanitize/${tool}_agent.log gpurun_out/sanitize/${tool}_env_replay.log
done
$T $SAN --tool synccheck --print-limit 50 --error-exitcode 0 python -m pytest $AGENT -m gpu -q -x > gpurun_out/sanitize/synccheck_agent.log 2>&1
tail -4 gpurun_out/sanitize/synccheck_agent.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitize/*.log
