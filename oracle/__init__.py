"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the Recovery RL hot path (env step, replay sampling, SAC + Q_risk
update) used ONLY as the checker: tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import it.  The product path
(recovery-rl_b200/) never imports anything from here and fails loudly if the CUDA
library is missing.

Pinning: Navigation1/2, replay and the SAC/Q_risk networks are pinned against the
reference's own code run in the build container (oracle/ref_harness/make_golden.py ->
tests/golden/*.npz).  Maze dynamics: PARITY UNPINNED -- the reference delegates them to the
closed MuJoCo 1.50 binary (env/maze.py:10,117; install.sh:13), which is absent here; the
restatement in oracle/envs.py follows env/maze.py + env/assets/simple_maze.xml and the
rules in SURVEY.md §8c, and the CUDA kernel is bit-exact against that restatement only.
"""
