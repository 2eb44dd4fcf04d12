"""Small runs of the paths bench.py's headline does not touch, for ncu launch lists / captures of their kernels:
a GridWorld collect on the tensor-core pair kernel (compact table, H = 128), an AlphaZero collect (k_mcts_*, k_az_finish),
evaluate (k_solve_step) and a deep-stack policy collect (k_forward_generic).  GPU box only."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import twisterl_b200 as tw  # noqa: E402
from helpers import synth_deep_state_dict  # noqa: E402
from parity import make_policies_general  # noqa: E402
from twisterl_b200 import collector as twc, nn as twn  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16x2w16"
tw.configure(device=0, precision=prec, seed=11)
eng = tw.default_engine()

# GridWorld 5x5 (625-row table -> 100 reachable rows, common width 128)
gsd = bench.synth_weights(obs_size=625, hidden=128)
gpol = bench.synth_policy(twn, gsd, 625)
genv = tw.env.GridWorld(5, 5, 64, 10)
gcol = twc.PPOCollector(65536, 0.995, 0.995, 32, engine=eng)
for _ in range(2):
    c = gcol.collect_device(genv, gpol)
print("gridworld records", c.n_records)

# AlphaZero on puzzle8
asd = bench.synth_weights(obs_size=81, hidden=256)
apol = bench.synth_policy(twn, asd, 81)
aenv = tw.env.Puzzle(3, 3, 8, 2, 256)
acol = twc.AZCollector(4096, 50, 1.41, 1, 32, engine=eng)
c = acol.collect_device(aenv, apol)
print("az records", c.n_records)

# the reference's AZ shape class (small batch): the persistent whole-search kernel
pcol = twc.AZCollector(512, 200, 1.41, 1, 32, engine=eng)
c = pcol.collect_device(aenv, apol)
print("az (persistent search) records", c.n_records)

# evaluate (single_solve loop) and MCTS-guided evaluate
print("evaluate", tw.collector.evaluate(aenv, apol, 4096, False, 4, 0, 0, 1.41, 1, 32))

# deep stacks -> k_forward_generic
dsd = synth_deep_state_dict(3, 256, 512, (256, 128), (64,), (32,), 4)
dpol, _ = make_policies_general(dsd, 256)
denv = tw.env.Puzzle(4, 4, 8, 2, 256)
dcol = twc.PPOCollector(4096, 0.995, 0.995, 32, engine=eng)
c = dcol.collect_device(denv, dpol)
print("deep-stack records", c.n_records)
eng.synchronize()
