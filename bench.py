#!/usr/bin/env python
"""bench.py -- rollout env-steps/s including the policy forward, puzzle15 PPO (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|f16x2]

One "step" = one PPOCollector.collect of `--episodes` (default 65 536) puzzle15 episodes per GPU at
difficulty 128 (<= 257 records each): reset, per-step policy forward + Gumbel-max sampling + env step +
trajectory write, GAE, compaction into the reference's CollectedData layout.  Under torchrun every rank
runs its own env shard (weak scaling); NCCL carries only the per-step weight broadcast and the stats
all-reduce.  `--impl reference` times the CPU restatement of the reference collector (oracle/) on the
host cores -- the Rust binary cannot be built in this image (no cargo).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_PER_STEP = {"puzzle15": 272_896, "puzzle8": 269_312, "gridworld": 145_152}
# what the tensor pipe executes per env-step: the embedding as a dense one-hot GEMM over the hi and the lo table, and the
# hidden layer as 3 (f16x2) or 2 (f16x2w16) split products
EXECUTED_TENSOR_FLOP_PER_STEP = {"f16x2": 2 * (2 * 256 * 512) + 3 * (2 * 512 * 256), "f16x2w16": 2 * (2 * 256 * 512) + 2 * (2 * 512 * 256)}
DTYPE = {"fp32": "f32",
         "f16x2": "f16x2 (tcgen05; every operand split into fp16 hi+lo, f32 accumulate; 1e-5-grade)",
         "f16x2w16": "f16x2w16 (tcgen05; table and activations split into fp16 hi+lo, common-layer weight one fp16 term, f32 accumulate; "
                     "7e-4 worst case on the shipped trained weights, bar 1e-3)"}
METRIC = "rollout env-steps/sec incl. policy fwd (puzzle15 PPO)"
UNIT = "env-steps/s"


def synth_weights(seed=0, obs_size=256, emb=512, hidden=256, n_act=4):
    g = np.random.default_rng(seed)
    f = lambda *s: (g.standard_normal(s) * 0.05).astype(np.float32)
    return {"embeddings.weight": f(emb, obs_size), "embeddings.bias": np.zeros(emb, np.float32),
            "common.0.weight": f(hidden, emb), "common.0.bias": np.zeros(hidden, np.float32),
            "action.0.weight": f(n_act, hidden), "action.0.bias": np.zeros(n_act, np.float32),
            "value.0.weight": f(1, hidden), "value.0.bias": np.zeros(1, np.float32)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16=d.get("bf16_tflops_sustained", 1410.7), hbm=d.get("hbm_gbs", 6546.2), src="measured")
    return dict(bf16=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self._nv = self._h = None
        try:                                   # NVML is initialised here, outside the timed region
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as exc:               # NVML missing: report it rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def run(self):
        nv, h = self._nv, self._h
        if nv is None:
            return
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        try:
            while not self._halt.is_set():     # the timed region can be as short as ~80 ms: sample every 5 ms
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                self._halt.wait(0.005)
        except Exception as exc:
            self.reasons.add(f"nvml_error:{type(exc).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------ reference arm ---
def oracle_policy(sd):
    from oracle import orc
    return orc.Policy.from_torch_state_dict(sd)


def cpu_collect_rate(sd, episodes, difficulty, threads, repeats=1):
    """records/s of the CPU restatement of PPOCollector::collect (oracle/twr_oracle.c)."""
    from oracle import orc
    spec = orc.puzzle_spec(4, 4, difficulty, 2, 256)
    pol = oracle_policy(sd)
    best = 0.0
    recs = 0
    for i in range(repeats):
        n, dt = orc.time_ppo_collect(spec, pol, episodes, 0.995, 0.995, 0x5EED5EED, i, threads)
        best = max(best, n / dt)
        recs = n
    return best, recs


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sd = synth_weights()
    episodes = args.ref_episodes
    for _ in range(args.warmup):
        cpu_collect_rate(sd, max(64, episodes // 8), args.difficulty, cores)
    t0 = time.perf_counter()
    total = 0
    from oracle import orc
    spec = orc.puzzle_spec(4, 4, args.difficulty, 2, 256)
    pol = oracle_policy(sd)
    for i in range(args.steps):
        n, _ = orc.time_ppo_collect(spec, pol, episodes, 0.995, 0.995, 0x5EED5EED, 100 + i, cores)
        total += n
    dt = time.perf_counter() - t0
    v = total / dt
    sample = f"{episodes} episodes/step of puzzle15 difficulty {args.difficulty} ({total} records in {dt:.2f}s)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"puzzle15 PPO rollout, {episodes} episodes/step (bounded sample of the 65536-env "
                               "config), CPU restatement of the reference Rust collector"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ------------------------------------------------------------------------ our arm ---
def run_ours(args, rank, local_rank, world):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import twisterl_b200 as tw
    from twisterl_b200 import _lib, collector as twc, nn as twn

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    eng = tw.Engine(device=local_rank, precision=args.precision, seed=0x5EED5EED, rank=rank, world=world,
                    stream=stream.cuda_stream)

    sd = synth_weights()
    obs_perms, act_perms = ([], [])
    if args.twists:      # {identity, main-diagonal transpose} twist set of SURVEY.md section 8a row T (BASELINE config 5)
        T = [(i % 4) * 4 + (i // 4) for i in range(16)]
        obs_perms = [list(range(256)), [T[i] * 16 + T[v] for i in range(16) for v in range(16)]]
        act_perms = [[0, 1, 2, 3], [1, 0, 3, 2]]
    pol = twn.Policy(twn.EmbeddingBag(sd["embeddings.weight"].T, sd["embeddings.bias"], True, [256], 0),
                     twn.Sequential([twn.Linear(sd["common.0.weight"].T.flatten(), sd["common.0.bias"], True)]),
                     twn.Sequential([twn.Linear(sd["action.0.weight"].T.flatten(), sd["action.0.bias"], False)]),
                     twn.Sequential([twn.Linear(sd["value.0.weight"].T.flatten(), sd["value.0.bias"], False)]),
                     obs_perms, act_perms)
    env = tw.env.Puzzle(4, 4, args.difficulty, 2, 256)
    col = twc.PPOCollector(args.episodes, 0.995, 0.995, 32, engine=eng)
    hpol = pol.device_handle(eng)
    L = _lib.load()
    nblob = int(L.twr_policy_blob_floats(hpol))
    dptr = C.c_void_p()
    _lib.check(L.twr_policy_blob_device_ptr(hpol, C.byref(dptr)))
    # trainer-side copy of the weights: what rank 0 broadcasts every iteration
    blob = torch.empty(nblob, dtype=torch.float32, device=dev)
    # initialise the broadcast source from the engine's blob (device-to-device through torch)
    blob.copy_(_torch_view(torch, dptr.value, nblob, dev))
    stats = torch.zeros(4, dtype=torch.float64, device=dev)

    prof = os.environ.get("TWISTERL_BENCH_PROFILE") and rank == 0
    phase = [0.0, 0.0, 0.0, 0.0]

    def one_step():
        t0 = time.perf_counter()
        if world > 1:
            dist.broadcast(blob, src=0)                              # per-iteration weight broadcast (NCCL)
        if prof:
            torch.cuda.synchronize(); t1 = time.perf_counter(); phase[0] += t1 - t0
        _lib.check(L.twr_policy_update_from_device(hpol, C.c_void_p(blob.data_ptr())))
        if prof:
            torch.cuda.synchronize(); t2 = time.perf_counter(); phase[1] += t2 - t1
        c = col.collect_device(env, pol)
        if prof:
            t3 = time.perf_counter(); phase[2] += t3 - t2
        if world > 1:
            stats.copy_(torch.tensor([c.num_episodes, c.successes, c.reward_sum, c.n_records], dtype=torch.float64))
            dist.all_reduce(stats)                                   # stats reduction (NCCL)
            stats.cpu()                                              # the reduced statistics are read on the host every
                                                                     # iteration (what a trainer logs); it also keeps the host
                                                                     # from queueing the next step's NCCL calls behind a peer
        if prof:
            torch.cuda.synchronize(); phase[3] += time.perf_counter() - t3
            print("[profile] broadcast %.3f update %.3f collect %.3f allreduce %.3f ms (cumulative)" % tuple(1e3 * x for x in phase), file=sys.stderr)
        return int(c.n_records)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    # ---- timed region: device-resident inputs, CUDA events on the launching stream
    eng.set_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    records, fwd_ms, fwd_launches = 0, 0.0, 0
    for _ in range(args.steps):
        records += one_step()
        f, _, n = eng.last_timing()
        fwd_ms += f; fwd_launches += n
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    eng.set_timing(False)
    agg = torch.tensor([ms, float(records), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, records, launches = float(mx[0]), float(sm[1]), float(sm[2])
    value = records / (ms * 1e-3)

    # ---- e2e: the C-ABI call with HOST buffers (weights H2D, CollectedData D2H inside the timed region)
    cap = int(L.twr_max_records(C.byref(tw.env.spec_from_env(env)), args.episodes))
    hb, arrs, holders = twc._host_buffers(cap, 16, 4, args.episodes, pinned=True, obs_u8=True)   # obs as u8 indices (obs_size 256)
    desc = pol.desc()
    spec = tw.env.spec_from_env(env)
    out = _lib.Collected()
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2 if args.warmup >= 2 else 1):
        _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), args.episodes, 0.995, 0.995,
                                          C.byref(hb), C.byref(out)))
    barrier()
    t0 = time.perf_counter()
    e2e_records = 0
    for _ in range(e2e_steps):
        _lib.check(L.twr_ppo_collect_host(eng._h, C.byref(spec), hpol, C.byref(desc), args.episodes, 0.995, 0.995,
                                          C.byref(hb), C.byref(out)))
        e2e_records += int(out.n_records)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e = torch.tensor([e2e_s, float(e2e_records)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = e2e.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = e2e.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_s, e2e_records = float(mx[0]), float(sm[1])
    rec_per_step = e2e_records / e2e_steps / world
    d2h = int(rec_per_step * sum(a.dtype.itemsize * int(np.prod(a.shape[1:])) for k, a in arrs.items() if k != "ep_len")
              + arrs["ep_len"].nbytes)
    h2d = nblob * 4

    if rank == 0:
        pk = peaks()
        flops = (records / world if world > 1 else records) * FLOP_PER_STEP["puzzle15"]
        # rank-0 forward time covers rank-0 records only
        r0_records = records / world
        achieved = r0_records * FLOP_PER_STEP["puzzle15"] / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE[args.precision],
            "data": "synthetic",
            "config": {"workload": f"examples/ppo_puzzle15_v1.json PPO rollout, {args.episodes} parallel envs per GPU, "
                                   f"difficulty {args.difficulty}, depth budget {2 * args.difficulty}, synthetic N(0,0.05^2) weights",
                       "episodes_per_gpu": args.episodes, "records_per_step": records / args.steps,
                       "l2": "working set per step (records + compacted output, > 1.5 GB at 65536 envs) exceeds the 126 MB L2; no flush needed",
                       "precision": args.precision, "twists": bool(args.twists)},
            "clocks": clocks,
            "e2e": {"value": e2e_records / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "twr_ppo_collect_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "k_forward_fp32" if args.precision == "fp32" else "k_forward_tc2",
                         "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                         "frac": (achieved / pk["bf16"]) if achieved else None,
                         # dram read+write bytes of one 32-step k_forward_tc2 launch over 65536 envs (profiles/r1f_tc2_summary.md;
                         # the records it writes -- 88 MB algorithmic, the rest still sits in L2 when the launch ends);
                         # steady-state launches cover 128 steps and move 4x as much
                         "traffic": 44_504_576 if (args.precision != "fp32" and args.episodes == 65536) else None,
                         "traffic_launch_steps": 32,
                         "peak_source": pk["src"] + " bf16_tflops_sustained",
                         "forward_ms_per_launch": fwd_ms / max(fwd_launches, 1), "forward_share_of_step": fwd_ms / ms,
                         "algorithmic_flop_per_env_step": FLOP_PER_STEP["puzzle15"]},
        }
        if args.precision != "fp32" and achieved:
            # what the tensor pipe executes for the fp32-grade result: the embedding as a dense one-hot GEMM (x2: table
            # hi/lo) and the hidden layer as 3 split products -- the tensor-pipe utilisation ncu reports follows this figure
            ex = EXECUTED_TENSOR_FLOP_PER_STEP[args.precision]
            line["roofline"]["executed_tensor_flop_per_env_step"] = ex
            line["roofline"]["executed"] = achieved * ex / FLOP_PER_STEP["puzzle15"]
            line["roofline"]["executed_frac"] = line["roofline"]["executed"] / pk["bf16"]
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate0, _ = cpu_collect_rate(sd, 256, args.difficulty, cores)
            episodes = int(min(65536, max(256, rate0 * 12 / (2 * args.difficulty + 1))))
            t0 = time.perf_counter()
            rate, recs = cpu_collect_rate(sd, episodes, args.difficulty, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{episodes} episodes of the same workload ({recs} records, "
                                              f"{time.perf_counter() - t0:.1f}s), C restatement of the Rust collector"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _torch_view(torch, ptr, n_floats, dev):
    """float32 torch view over an engine-owned device buffer (CUDA array interface)."""
    class _Holder:
        __cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Holder(), device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("TWISTERL_B200_PRECISION", "f16x2w16"), choices=["fp32", "f16x2", "f16x2w16"])
    ap.add_argument("--episodes", type=int, default=65536)
    ap.add_argument("--difficulty", type=int, default=128)
    ap.add_argument("--ref-episodes", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--twists", action="store_true", help="enable the {identity, transpose} twist set (BASELINE config 5)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
