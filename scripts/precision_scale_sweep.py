import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import twisterl_b200 as tw
from helpers import synth_state_dict, scramble_states, obs_from_states
from parity import make_policies
from twisterl_b200 import _lib
from twisterl_b200.env import EnvBatch
from twisterl_b200.nn import forward_batch
from oracle import orc

def f64_ref(sd, obs):
    E, W = sd["embeddings.weight"].T.astype(np.float64), sd["common.0.weight"].T.astype(np.float64)
    h1 = np.maximum(E[obs].sum(1) + sd["embeddings.bias"], 0)
    h2 = np.maximum(h1 @ W + sd["common.0.bias"], 0)
    return (h2 @ sd["action.0.weight"].T.astype(np.float64) + sd["action.0.bias"], (h2 @ sd["value.0.weight"].T.astype(np.float64) + sd["value.0.bias"])[:, 0])
err = lambda a, r: float((np.abs(a - r) / np.maximum(1.0, np.abs(r))).max())
st = scramble_states(np.random.default_rng(1), 4096, 4, 4, 200)
obs = obs_from_states(st)
spec = _lib.EnvSpec(0, 4, 4, 1, 2, 256)
for se, sw, sb in [(1, 1, 0), (8, 1, 0), (1, 8, 0), (8, 8, 0), (30, 4, 0), (0.05, 1, 0), (1, 0.05, 0), (1, 1, 3.0), (200, 1, 0)]:
    sd = synth_state_dict(3, 256, 512, 256, 4)
    sd["embeddings.weight"] = (sd["embeddings.weight"] * se).astype(np.float32)
    sd["common.0.weight"] = (sd["common.0.weight"] * sw).astype(np.float32)
    sd["embeddings.bias"] = (sd["embeddings.bias"] + sb).astype(np.float32)
    rl, rv = f64_ref(sd, obs)
    line = f"emb x{se:<5} W x{sw:<5} bias+{sb}: max|logit| {np.abs(rl).max():9.2f} h1max {np.maximum(sd['embeddings.weight'].T[obs].sum(1)+sd['embeddings.bias'],0).max():8.1f} |"
    for prec in ("f16x2", "f16x2w16", "f16f8c"):
        eng = tw.Engine(device=0, precision=prec, seed=1)
        pol, _ = make_policies(sd, 256)
        b = EnvBatch(spec, len(st), eng); b.set_state(st)
        l, v = forward_batch(eng, pol, b)
        line += f" {prec} {err(l, rl):.2e}/{err(v, rv):.2e}"
        pol.release(); eng.close()
    print(line, flush=True)
