"""Device time of a PPO collect against the length of the persistent launches (TWISTERL_B200_CHUNK), for workloads whose
episodes end early.  usage: chunk_sweep.py gridworld|puzzle8 [envs]"""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import twisterl_b200 as tw
from helpers import synth_state_dict
from parity import make_policies

kind = sys.argv[1] if len(sys.argv) > 1 else "gridworld"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
if kind == "gridworld":
    sd, obs, env = synth_state_dict(0, 625, 512, 128, 4), 625, tw.env.GridWorld(5, 5, 64, 10)
else:
    sd, obs, env = synth_state_dict(0, 81, 512, 256, 4), 81, tw.env.Puzzle(3, 3, 32, 2, 64)
eng = tw.Engine(device=0, precision="f16f8c", seed=1)
pol, _ = make_policies(sd, obs)
col = tw.collector.PPOCollector(E, 0.995, 0.995, 1, engine=eng)
for chunk in (0, 1, 2, 3, 4, 6, 9, 12, 17, 33):
    if chunk: os.environ["TWISTERL_B200_CHUNK"] = str(chunk)
    else: os.environ.pop("TWISTERL_B200_CHUNK", None)
    ts, recs = [], 0
    for rep in range(6):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record()
        d = col.collect_device(env, pol)
        t1.record(); torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1)); recs = d.n_records
    t = float(np.median(ts[2:]))
    print(f"{kind} {E} envs, chunk {chunk or 'default':>7}: {t:7.3f} ms, {recs} records, {recs / t / 1e6:.3f}e9 env-steps/s", flush=True)
