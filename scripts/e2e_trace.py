"""One traced twr_ppo_collect_host call per sub-batch split (TWISTERL_B200_E2E_TRACE milestones on stderr)."""
import ctypes as C, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
import twisterl_b200 as tw
from twisterl_b200 import _lib, collector as twc, nn as twn
E = 65536
sd = bench.synth_weights()
eng = tw.Engine(device=0, precision="f16x2w16", seed=0x5EED5EED)
pol = twn.Policy(twn.EmbeddingBag(sd["embeddings.weight"].T, sd["embeddings.bias"], True, [256], 0),
                 twn.Sequential([twn.Linear(sd["common.0.weight"].T.flatten(), sd["common.0.bias"], True)]),
                 twn.Sequential([twn.Linear(sd["action.0.weight"].T.flatten(), sd["action.0.bias"], False)]),
                 twn.Sequential([twn.Linear(sd["value.0.weight"].T.flatten(), sd["value.0.bias"], False)]), [], [])
env = tw.env.Puzzle(4, 4, 128, 2, 256)
L = _lib.load(); spec = tw.env.spec_from_env(env)
cap = int(L.twr_max_records(C.byref(spec), E))
hb, arrs, keep = twc._host_buffers(cap, 16, 4, E, pinned=True, obs_u8=True)
desc = pol.desc(); hpol = pol.device_handle(eng); out = _lib.Collected()
for parts in (None, "37888,27648", "18944,27648,18944"):
    for nopack in (0, 1):
        os.environ.pop("TWISTERL_B200_E2E_PARTS", None); os.environ.pop("TWISTERL_B200_E2E_NOPACK", None); os.environ.pop("TWISTERL_B200_E2E_TRACE", None)
        if parts: os.environ["TWISTERL_B200_E2E_PARTS"] = parts
        if nopack: os.environ["TWISTERL_B200_E2E_NOPACK"] = "1"
        for _ in range(3):
            _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), E, 0.995, 0.995, C.byref(hb), C.byref(out)))
        os.environ["TWISTERL_B200_E2E_TRACE"] = "1"
        print(f"--- parts={parts} nopack={nopack}", file=sys.stderr, flush=True)
        _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), E, 0.995, 0.995, C.byref(hb), C.byref(out)))
keep = None
