"""Pipeline trace of the first persistent launch of a full-size collect (TWISTERL_B200_TRACE=0), per precision."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["TWISTERL_B200_TRACE"] = "0"
import twisterl_b200 as tw
from helpers import synth_state_dict
from parity import make_policies

sd = synth_state_dict(0, 256, 512, 256, 4)
for prec in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["f16x2w16", "f16f8c"]):
    eng = tw.Engine(device=0, precision=prec, seed=1)
    pol, _ = make_policies(sd, 256)
    env = tw.env.Puzzle(4, 4, 128, 2, 256)
    col = tw.collector.PPOCollector(65536, 0.995, 0.995, 1, engine=eng)
    for rep in range(3):
        print(f"== {prec} collect {rep}", file=sys.stderr, flush=True)
        col.collect_device(env, pol)
    pol.release(); eng.close()
