"""Tiny runs of the round-2 kernels for `compute-sanitizer --tool memcheck` (GPU box only): GridWorld collect on the pair
kernel (compact table, H = 128), a puzzle15 collect in both tensor-core precisions, the packed host collect, and an
AlphaZero collect small enough for the persistent search kernel."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import twisterl_b200 as tw  # noqa: E402
from twisterl_b200 import _lib, collector as twc, nn as twn  # noqa: E402

for prec in ("f16x2w16", "f16x2"):
    eng = tw.Engine(device=0, precision=prec, seed=3)
    pol = bench.synth_policy(twn, bench.synth_weights(), 256, *bench.puzzle15_twists())
    c = twc.PPOCollector(700, 0.995, 0.995, 1, engine=eng).collect_device(tw.env.Puzzle(4, 4, 6, 2, 256), pol)
    print(prec, "puzzle15 records", c.n_records)
    gpol = bench.synth_policy(twn, bench.synth_weights(obs_size=625, hidden=128), 625)
    c = twc.PPOCollector(900, 0.995, 0.995, 1, engine=eng).collect_device(tw.env.GridWorld(5, 5, 64, 10), gpol)
    print(prec, "gridworld records", c.n_records)
    apol = bench.synth_policy(twn, bench.synth_weights(obs_size=81, hidden=256), 81)
    c = twc.AZCollector(70, 20, 1.41, 1, 1, engine=eng).collect_device(tw.env.Puzzle(3, 3, 4, 2, 256), apol)
    print(prec, "az records", c.n_records)
    # packed, pipelined host collect
    import os
    os.environ["TWISTERL_B200_E2E_SPLIT"] = "3"
    env = tw.env.Puzzle(4, 4, 5, 2, 256)
    L = _lib.load(); spec = tw.env.spec_from_env(env)
    cap = int(L.twr_max_records(C.byref(spec), 600))
    hb, arrs, keep = twc._host_buffers(cap, 16, 4, 600, pinned=True, obs_u8=True)
    out = _lib.Collected()
    _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), pol.device_handle(eng), None, 600, 0.995, 0.995, C.byref(hb), C.byref(out)))
    print(prec, "host collect records", out.n_records)
    os.environ.pop("TWISTERL_B200_E2E_SPLIT")
    keep = None
    pol.release(); gpol.release(); apol.release(); eng.close()
print("done")
