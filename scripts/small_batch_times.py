"""Device time of one PPO collect (puzzle15, difficulty 128) at sub-batch sizes, and the host-collect timeline of a few
splits.  GPU box only."""
import ctypes as C, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
import twisterl_b200 as tw
from twisterl_b200 import _lib, collector as twc, nn as twn
sd = bench.synth_weights()
eng = tw.Engine(device=0, precision=os.environ.get("TWISTERL_B200_PRECISION", "f16f8c"), seed=0x5EED5EED)
pol = bench.synth_policy(twn, sd, 256)
env = tw.env.Puzzle(4, 4, 128, 2, 256)
eng.set_timing(True)
for E in (8704, 18944, 27648, 37888, 46592, 56832, 65536):
    col = twc.PPOCollector(E, 0.995, 0.995, 32, engine=eng)
    for _ in range(3):
        c = col.collect_device(env, pol)
    f, t, n = eng.last_timing()
    print(f"{E:6d} envs: {t:.2f} ms total, {f:.2f} ms in {n} forward launches, {c.n_records / t / 1e6:.1f}e9 env-steps/s", flush=True)
eng.set_timing(False)
L = _lib.load(); spec = tw.env.spec_from_env(env)
E = 65536
cap = int(L.twr_max_records(C.byref(spec), E))
hb, arrs, keep = twc._host_buffers(cap, 16, 4, E, pinned=True, obs_u8=True)
desc = pol.desc(); hpol = pol.device_handle(eng); out = _lib.Collected()
for parts in ("37888,27648", "37888,18944,8704", "18944,37888,8704", "18944,27648,18944", "18944,18944,18944,8704", "8704,37888,18944", "18944,46592"):
    os.environ["TWISTERL_B200_E2E_PARTS"] = parts
    os.environ.pop("TWISTERL_B200_E2E_TRACE", None)
    for _ in range(3):
        _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), E, 0.995, 0.995, C.byref(hb), C.byref(out)))
    t0 = time.perf_counter()
    for _ in range(5):
        _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), E, 0.995, 0.995, C.byref(hb), C.byref(out)))
    print(f"parts {parts}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call", flush=True)
    os.environ["TWISTERL_B200_E2E_TRACE"] = "1"
    _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), E, 0.995, 0.995, C.byref(hb), C.byref(out)))
keep = None
