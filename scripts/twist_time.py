import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import twisterl_b200 as tw
from helpers import synth_state_dict, transpose_twists
from parity import make_policies
sd = synth_state_dict(0, 256, 512, 256, 4)
eng = tw.Engine(device=0, precision="f16f8c", seed=1)
for tw_on in (False, True):
    pol, _ = make_policies(sd, 256, *(transpose_twists(4) if tw_on else ((), ())))
    env = tw.env.Puzzle(4, 4, 128, 2, 256)
    col = tw.collector.PPOCollector(65536, 0.995, 0.995, 1, engine=eng)
    ts = []
    for rep in range(6):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record(); d = col.collect_device(env, pol); t1.record(); torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1))
    print(f"twists {tw_on}: {np.median(ts[2:]):.2f} ms ({d.n_records / np.median(ts[2:]) / 1e6:.3f}e9 env-steps/s)", flush=True)
    pol.release()
