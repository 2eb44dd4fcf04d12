"""Which stage bounds the fused persistent forward: device time of a full collect (65 536 puzzle15 envs, difficulty 128)
per precision with parts of the kernel switched off (ForwardArgs::dbg_flags via TWISTERL_B200_TC_FLAGS: 1 = no epilogue-1
conversion, 2 = no operand TMA traffic, 4 = no GEMM2 MMAs, 8 = no GEMM1 MMAs).  Results of a run with flags are garbage;
only the time is meaningful."""
import os
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import twisterl_b200 as tw
from helpers import synth_state_dict
from parity import make_policies

precisions = sys.argv[1].split(",") if len(sys.argv) > 1 else ["f16x2w16", "f16f8c"]
flag_sets = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 4, 8, 12, 13]
sd = synth_state_dict(0, 256, 512, 256, 4)
for prec in precisions:
    for flags in flag_sets:
        os.environ["TWISTERL_B200_TC_FLAGS"] = str(flags)
        eng = tw.Engine(device=0, precision=prec, seed=1)
        pol, _ = make_policies(sd, 256)
        env = tw.env.Puzzle(4, 4, 128, 2, 256)
        col = tw.collector.PPOCollector(65536, 0.995, 0.995, 1, engine=eng)
        ts = []
        for rep in range(6):
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); t0.record()
            d = col.collect_device(env, pol) if hasattr(col, "collect_device") else col.collect_torch(env, pol)
            t1.record(); torch.cuda.synchronize()
            ts.append(t0.elapsed_time(t1))
        print(f"{prec:9s} flags {flags:2d}: {np.median(ts[2:]):7.2f} ms  (all {', '.join('%.2f' % t for t in ts)})", flush=True)
        pol.release(); eng.close()
