"""C++ host side above the C ABI (include/twisterl_b200.hpp: the compiled-language mirror of the reference's
Collector trait / CollectedData / nn types).  CPU: it compiles and links as C++17 and fails loudly without a device.
GPU: a collect through it is byte-identical to the Python mirror's collect on the same Philox streams."""
import os
import shutil
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from suite_loader import suite_precision

from helpers import synth_state_dict

ROOT = Path(__file__).resolve().parent.parent
PRECISION = suite_precision(globals())


def _build(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    libdir = ROOT / "twisterl_b200" / "lib"
    exe = tmp_path / "collect_check"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
                        str(ROOT / "tests" / "cpp" / "collect_check.cpp"), f"-L{libdir}", "-ltwisterl_b200",
                        f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _weights_file(tmp_path, sd):
    parts = [sd["embeddings.weight"].T, sd["embeddings.bias"], sd["common.0.weight"].T, sd["common.0.bias"],
             sd["action.0.weight"].T, sd["action.0.bias"], sd["value.0.weight"].T, sd["value.0.bias"]]
    p = tmp_path / "weights.bin"
    p.write_bytes(b"".join(np.ascontiguousarray(a, dtype=np.float32).tobytes() for a in parts))
    return p


def test_cpp_host_builds_and_fails_loudly_without_a_device(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present; the GPU test covers the run")
    w = _weights_file(tmp_path, synth_state_dict(4, 256, 512, 256, 4))
    r = subprocess.run([str(exe), str(w), str(tmp_path / "out.bin"), "1", "7", "3", "16", "4"], capture_output=True, text=True)
    assert r.returncode == 1 and "error:" in r.stderr          # twisterl::Error from Engine's constructor: no CPU fallback


@pytest.mark.gpu
def test_cpp_collect_matches_python_mirror(tmp_path):
    import twisterl_b200 as tw
    from parity import make_policies
    exe = _build(tmp_path)
    sd = synth_state_dict(4, 256, 512, 256, 4)
    w = _weights_file(tmp_path, sd)
    seed, cid, E, diff = 0xABCDEF, 5, 300, 6
    out = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(w), str(out), str({"fp32": 0, "f16x2": 1, "f16x2w16": 2, "f16f8c": 3}[PRECISION]), str(seed), str(cid), str(E), str(diff)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    R = struct.unpack_from("<q", raw, 0)[0]
    rec = np.frombuffer(raw, dtype=np.dtype([("obs", "<i4", 16), ("logits", "<f4", 4), ("action", "<i4"), ("perm", "<i4"),
                                            ("value", "<f4"), ("reward", "<f4"), ("adv", "<f4"), ("ret", "<f4")]), count=R, offset=8)
    ev = struct.unpack_from("<ff", raw, 8 + R * rec.dtype.itemsize)
    eng = tw.Engine(device=0, precision=PRECISION, seed=seed)
    pol, _ = make_policies(sd, 256)
    env = tw.env.Puzzle(4, 4, diff, 2, 256)
    eng.set_collect_id(cid)
    d = tw.collector.PPOCollector(E, 0.995, 0.995, 32, engine=eng).collect(env, pol)
    assert R == len(d.values_array)
    assert np.array_equal(rec["obs"], d.obs_array.astype(np.int32)) and np.array_equal(rec["action"], d.actions_array.astype(np.int32))
    assert np.array_equal(rec["logits"], d.logits_array) and np.array_equal(rec["value"], d.values_array)
    assert np.array_equal(rec["reward"], d.rewards_array) and np.array_equal(rec["perm"], d.perms_array.astype(np.int32))
    assert np.array_equal(rec["adv"], d.additional_array("advs")) and np.array_equal(rec["ret"], d.additional_array("rets"))
    assert 0.0 <= ev[0] <= 1.0
    pol.release(); eng.close()
