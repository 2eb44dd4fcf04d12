"""Does a concurrent copy slow a rollout down?  One PPO collect (puzzle15, difficulty 128) of a few batch sizes alone and
while a 400 MB copy (D2H to pinned memory / D2D / H2D) runs on another stream.  GPU box only."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
import twisterl_b200 as tw
from twisterl_b200 import collector as twc, nn as twn

sd = bench.synth_weights()
eng = tw.Engine(device=0, precision="f16x2w16", seed=0x5EED5EED)
pol = bench.synth_policy(twn, sd, 256)
env = tw.env.Puzzle(4, 4, 128, 2, 256)
N = 400 << 20
dev_a = torch.empty(N, dtype=torch.uint8, device="cuda"); dev_b = torch.empty(N, dtype=torch.uint8, device="cuda")
host = torch.empty(N, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()
eng.set_timing(True)

def copy(kind):
    with torch.cuda.stream(side):
        for _ in range(3):                     # ~3 x 7 ms: covers the whole collect
            if kind == "d2h": host.copy_(dev_a, non_blocking=True)
            elif kind == "h2d": dev_a.copy_(host, non_blocking=True)
            elif kind == "d2d": dev_b.copy_(dev_a, non_blocking=True)

for E in (18944, 27648, 37888, 65536):
    col = twc.PPOCollector(E, 0.995, 0.995, 32, engine=eng)
    for _ in range(3):
        col.collect_device(env, pol)
    row = []
    for kind in ("none", "d2h", "h2d", "d2d"):
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            if kind != "none":
                copy(kind)
            col.collect_device(env, pol)
            ts.append(eng.last_timing()[1])
            torch.cuda.synchronize()
        row.append(f"{kind} {min(ts):.2f}")
    print(f"{E:6d} envs: collect ms with concurrent copy: " + " | ".join(row), flush=True)
