"""Debug helper: run the f16x2 tensor-core forward on the golden states and print the error pattern."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import twisterl_b200 as tw
from helpers import trained15
from oracle import orc
from parity import make_policies
from twisterl_b200 import _lib
from twisterl_b200.env import EnvBatch
from twisterl_b200.nn import forward_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
z, sd = trained15()
pol, opol = make_policies(sd, 256)
eng = tw.Engine(device=0, precision="f16x2", seed=1)
st = np.tile(z["states"], (max(1, n // 512 + 1), 1))[:n]
b = EnvBatch(_lib.EnvSpec(0, 4, 4, 1, 2, 256), n, eng)
b.set_state(st)
logits, values = forward_batch(eng, pol, b)
rl = np.tile(z["logits"], (max(1, n // 512 + 1), 1))[:n]; rv = np.tile(z["values"], max(1, n // 512 + 1))[:n]
el = np.abs(logits - rl) / np.maximum(1, np.abs(rl)); ev = np.abs(values - rv) / np.maximum(1, np.abs(rv))
print("n", n, "max logit err", el.max(), "max value err", ev.max())
print("first rows gpu", logits[:3], values[:3]); print("first rows ref", rl[:3], rv[:3])
bad = np.where(el.max(axis=1) > 1e-3)[0]
print("bad rows:", len(bad), bad[:40])
