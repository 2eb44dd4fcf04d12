"""Per-CTA pipeline wait counters of k_forward_tc (twr_debug_forward_profile)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import twisterl_b200 as tw
from helpers import synth_state_dict, scramble_states
from parity import make_policies
from twisterl_b200 import _lib
from twisterl_b200.env import EnvBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sd = synth_state_dict(0, 256, 512, 256, 4)
pol, _ = make_policies(sd, 256)
eng = tw.Engine(device=0, precision="f16x2", seed=1)
b = EnvBatch(_lib.EnvSpec(0, 4, 4, 64, 2, 256), n, eng)
b.reset()
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
c = np.zeros((148, 16), dtype=np.int64)
for rep in range(3):
    _lib.check(_lib.load().twr_debug_forward_profile(eng._h, pol.device_handle(eng), b._h, _lib.ptr(c), 148, flags))
print("flags", flags)
names = {0: "mma total", 1: "mma wait slot(TMA)", 2: "mma wait a1_full", 3: "mma wait a2_full", 4: "mma wait d2_empty",
         5: "tiles", 6: "mma wait peer slot", 8: "producer wait empty", 9: "epi total", 10: "epi wait d1_full", 11: "epi wait d2_full", 12: "epi wait a1_empty"}
for k, v in names.items():
    col = c[:, k]
    print(f"{v:22s} cta0 {col[0]:9d}  mean {col.mean():11.1f}  min {col.min():9d}  max {col.max():9d}")
