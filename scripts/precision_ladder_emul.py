"""CPU emulation of the tensor-core forward's precision ladder on the trained puzzle15 weights.

Every operand is rounded as the kernel would round it (fp16 hi / hi+lo split, or bf16), products are accumulated in
fp32 (numpy float32 matmul; the TMEM accumulator is fp32), heads run in fp32 like the kernel's CUDA-core heads.
Error measure = the parity tests' own: max |d| / max(1, |ref|) against the reference torch BasicPolicy outputs
stored in tests/golden/policy15_trained.npz.  Prints one row per (GEMM1 passes, GEMM2 passes) combination."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from helpers import trained15, obs_from_states

def bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)

def split(x, fmt):
    if fmt == "f16":
        hi = x.astype(np.float16).astype(np.float32)
        lo = (x - hi).astype(np.float16).astype(np.float32)
    else:
        hi = bf16(x); lo = bf16(x - hi)
    return hi, lo

def run(fmt, g1, g2, sd, obs):
    E = sd["embeddings.weight"].T.astype(np.float32)          # [obs, 512]
    Ehi, Elo = split(E, fmt)
    W = sd["common.0.weight"].T.astype(np.float32)            # [512, 256]
    Whi, Wlo = split(W, fmt)
    h1 = np.zeros((obs.shape[0], E.shape[1]), np.float32)
    for k in range(obs.shape[1]):
        h1 += Ehi[obs[:, k]]
        if g1 == 2:
            h1 += Elo[obs[:, k]]
    h1 = np.maximum(h1 + sd["embeddings.bias"], 0).astype(np.float32)
    ahi, alo = split(h1, fmt)
    acc = ahi @ Whi
    if g2 >= 2: acc = acc + alo @ Whi          # activation fully resolved, weight rounded
    if g2 >= 3: acc = acc + ahi @ Wlo
    h2 = np.maximum(acc + sd["common.0.bias"], 0).astype(np.float32)
    logits = h2 @ sd["action.0.weight"].T + sd["action.0.bias"]
    values = (h2 @ sd["value.0.weight"].T + sd["value.0.bias"])[:, 0]
    return logits, values

def err(a, ref):
    return float((np.abs(a - ref) / np.maximum(1.0, np.abs(ref))).max())

if __name__ == "__main__":
    z, sd = trained15()
    obs = obs_from_states(z["states"])
    print("max|logit| %.1f  max|value| %.2f" % (np.abs(z["logits"]).max(), np.abs(z["values"]).max()))
    for fmt in ("f16", "bf16"):
        for g1 in (1, 2):
            for g2 in (1, 2, 3):
                l, v = run(fmt, g1, g2, sd, obs)
                flop = 2 * 256 * 512 * g1 + 2 * 512 * 256 * g2
                print(f"{fmt:5s} G1x{g1} G2x{g2}  executed {flop:8d}  logits {err(l, z['logits']):.2e}  values {err(v, z['values']):.2e}")
