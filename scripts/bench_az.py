"""AlphaZero collection throughput (BASELINE config 4): puzzle8, batched MCTS, GPU vs the CPU oracle."""
import os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import twisterl_b200 as tw
from helpers import synth_state_dict
from oracle import orc
from parity import make_policies

E = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 100
diff = int(sys.argv[3]) if len(sys.argv) > 3 else 8
prec = os.environ.get("TWISTERL_B200_PRECISION", "f16x2")
sd = synth_state_dict(8, 81, 512, 256, 4)
pol, opol = make_policies(sd, 81)
eng = tw.Engine(device=0, precision=prec, seed=7)
env = tw.env.Puzzle(3, 3, diff, 2, 256)
col = tw.collector.AZCollector(E, sims, 1.41, 1, 32, engine=eng)
col.collect(env, pol)
times = []
for _ in range(int(os.environ.get("REPEATS", "5"))):
    t0 = time.perf_counter(); d = col.collect_device(env, pol); times.append(time.perf_counter() - t0)
dt = min(times)
R = int(d.n_records)
print("collect seconds:", " ".join(f"{t:.3f}" for t in times))
print(f"GPU  {prec}: {E} episodes x {sims} sims, difficulty {diff}: {R} records in {dt:.3f}s -> {R/dt:.1f} records/s, "
      f"{R*(sims+1)/dt:.3e} leaf evals/s, launches {eng.launch_count()}")
Ec = max(8, E // 32)
t0 = time.perf_counter(); oc = orc.az_collect(orc.puzzle_spec(3, 3, diff, 2, 256), opol, Ec, sims, 1.41, 1, seed=7); dt = time.perf_counter() - t0
print(f"CPU oracle (1 thread): {Ec} episodes: {oc['n_records']} records in {dt:.3f}s -> {oc['n_records']/dt:.1f} records/s "
      f"(x{os.cpu_count()} cores ~ {oc['n_records']/dt*os.cpu_count():.1f})")
