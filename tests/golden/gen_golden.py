"""Generate the committed golden fixtures from the reference itself.

Run ONLY in the build container (needs /root/reference); the fixtures it writes are what the
tests read, so nothing at test time touches /root/reference.

    python tests/golden/gen_golden.py

Outputs (tests/golden/):
  puzzle_replays.json   -- boards printed in the reference notebooks examples/puzzle.ipynb and
                           examples/hub_puzzle_model.ipynb (start state, action list, every
                           intermediate board) and examples/grid_world/game.ipynb
  policy15_trained.npz  -- weights of examples/ppo_puzzle15_v1.pt + logits/values produced by the
                           reference's torch BasicPolicy (src/twisterl/nn/policy.py) on seeded states,
                           without twists and with the {identity, transpose} twist set
  policy_synth.npz      -- same for seeded synthetic weights on the puzzle8 / grid_world shapes
                           (weights are regenerated from the seed at test time, only I/O is stored)
  policy_deep.npz       -- the reference's torch BasicPolicy with deeper stacks (common_layers=(256,128),
                           policy_layers=(64,), value_layers=(32,)), with and without twists, seeded synthetic weights
  policy_conv1d.npz     -- the reference's torch Conv1dPolicy (both conv_dim values, with and without twists) on
                           seeded synthetic weights, puzzle15 shape
"""
import json
import re
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

# ---- the reference's Python half imports `twisterl.twisterl` (the Rust extension); a stub is enough
sys.path.insert(0, str(REF / "src"))
stub = types.ModuleType("twisterl.twisterl")
for sub in ("nn", "env", "collector"):
    m = types.ModuleType(f"twisterl.twisterl.{sub}")
    setattr(stub, sub, m)
stub.env.Puzzle = object
stub.env.PyEnv = object
sys.modules["twisterl.twisterl"] = stub
import twisterl  # noqa: E402

twisterl.twisterl = stub
from twisterl.nn.policy import BasicPolicy, Conv1dPolicy  # noqa: E402

sys.path.insert(0, str(OUT.parent))
from helpers import scramble_states, synth_conv_state_dict, synth_deep_state_dict, synth_state_dict, transpose_twists  # noqa: E402


def parse_boards(text, names):
    """Notebook replay output -> (labels, boards)."""
    labels, boards, cur = [], [], []
    for line in text.splitlines():
        line = line.strip()
        if line.endswith(":") and not line.startswith("|"):
            labels.append(line[:-1])
        elif line.startswith("|"):
            cur.append([int(c) if c.strip() else 0 for c in line.strip("|").split("|")])
        elif cur:
            boards.append([v for row in cur for v in row])
            cur = []
    if cur:
        boards.append([v for row in cur for v in row])
    acts = [names.index(l) for l in labels[1:]]
    return acts, boards


def notebook_replays():
    out = {}
    for key, nb, names, cell in (
        ("puzzle8_35", "examples/puzzle.ipynb", ["left", "up", "right", "down"], 11),
        ("puzzle8_123", "examples/hub_puzzle_model.ipynb", ["left", "up", "right", "down"], 13),
        # NB: game.ipynb prints labels from a mis-ordered list ["up","down","right","left"]; the
        # action *ids* (what we replay) are recovered through that same list.
        ("gridworld_4", "examples/grid_world/game.ipynb", ["up", "down", "right", "left"], 12),
    ):
        d = json.load(open(REF / nb))
        text = "".join(o.get("text", "") if isinstance(o.get("text", ""), str) else "".join(o["text"])
                       for o in d["cells"][cell]["outputs"])
        acts, boards = parse_boards(text, names)
        assert len(boards) == len(acts) + 1, (key, len(boards), len(acts))
        out[key] = {"source": nb, "start": boards[0], "actions": acts, "boards": boards}
    # cross-check with the action list the hub notebook prints explicitly
    d = json.load(open(REF / "examples/hub_puzzle_model.ipynb"))
    printed = json.loads("".join(d["cells"][11]["outputs"][0]["text"]))
    assert printed == out["puzzle8_123"]["actions"]
    assert len(out["puzzle8_35"]["actions"]) == 35 and len(printed) == 123
    return out




def gridworld_states(rng, n, w, h):
    N = w * h
    states = np.zeros((n, N), dtype=np.uint8)
    for k in range(n):
        a, g, t = rng.choice(N, size=3, replace=False)
        states[k, g] = 2; states[k, t] = 3; states[k, a] = 1
    return states


def one_hot(states):
    n, N = states.shape
    x = np.zeros((n, N * N), dtype=np.float32)
    for i in range(N):
        x[np.arange(n), i * N + states[:, i].astype(np.int64)] = 1.0
    return torch.from_numpy(x)






def run_reference_policy(sd, N, emb, hidden, x, obs_perms=(), act_perms=(), perm_idx=None):
    pol = BasicPolicy([N, N], 4, emb, common_layers=(hidden,), obs_perms=obs_perms, act_perms=act_perms,
                      device="cpu")
    pol.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    pol.eval()
    with torch.no_grad():
        if perm_idx is None:
            l, v = pol(x)
        else:
            l, v = pol(x, perm_indices=torch.as_tensor(perm_idx))
    return l.numpy().astype(np.float32), v.numpy().astype(np.float32).reshape(-1)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    (OUT / "puzzle_replays.json").write_text(json.dumps(notebook_replays()))

    # ---- trained puzzle15 weights
    sd = torch.load(REF / "examples/ppo_puzzle15_v1.pt", weights_only=True, map_location="cpu")
    sd = {k: v.numpy().astype(np.float32) for k, v in sd.items()}
    rng = np.random.default_rng(15)
    states = scramble_states(rng, 512, 4, 4, 128)
    states[0] = np.arange(16)  # solved state: SURVEY 8c known answer
    x = one_hot(states)
    logits, values = run_reference_policy(sd, 16, 512, 256, x)
    assert np.allclose(logits[0], [-4.4116335, -5.16946, -2.0670972, 0.8337202], atol=1e-5)
    obs_perms, act_perms = transpose_twists(4)
    perm_idx = rng.integers(0, 2, size=len(states)).astype(np.int64)
    tl, tv = run_reference_policy(sd, 16, 512, 256, x, obs_perms, act_perms, perm_idx)
    np.savez(OUT / "policy15_trained.npz", states=states, logits=logits, values=values,
             twist_perm_idx=perm_idx.astype(np.int32), twist_logits=tl, twist_values=tv,
             **{"w." + k: v for k, v in sd.items()})

    # ---- synthetic weights on the other two config shapes
    out = {}
    for name, seed, N, hidden, gen in (("puzzle8", 8, 9, 256, lambda r, n: scramble_states(r, n, 3, 3, 32)),
                                       ("gridworld", 25, 25, 128, lambda r, n: gridworld_states(r, n, 5, 5))):
        rng = np.random.default_rng(seed)
        st = gen(rng, 256)
        ssd = synth_state_dict(seed, N * N, 512, hidden, 4)
        l, v = run_reference_policy(ssd, N, 512, hidden, one_hot(st))
        out[f"{name}.states"] = st; out[f"{name}.logits"] = l; out[f"{name}.values"] = v
        out[f"{name}.seed"] = np.int64(seed); out[f"{name}.hidden"] = np.int64(hidden)
    np.savez(OUT / "policy_synth.npz", **out)

    # ---- Conv1dPolicy (2-D EmbeddingBag path, nn/layers.rs:63-77) on the puzzle15 shape
    out = {}
    rng = np.random.default_rng(151)
    st = scramble_states(rng, 256, 4, 4, 64)
    x = one_hot(st)
    obs_perms, act_perms = transpose_twists(4)
    perm_idx = rng.integers(0, 2, size=len(st)).astype(np.int64)
    out["states"] = st; out["twist_perm_idx"] = perm_idx.astype(np.int32)
    for conv_dim in (0, 1):
        seed = 1510 + conv_dim
        csd = synth_conv_state_dict(seed, 16, 32, 512, 256, 4)
        for tag, perms in (("plain", ((), ())), ("twist", (obs_perms, act_perms))):
            pol = Conv1dPolicy([16, 16], 4, 512, conv_dim=conv_dim, common_layers=(256,), obs_perms=perms[0],
                               act_perms=perms[1])
            with torch.no_grad():
                pol.conv_layer.weight.copy_(torch.as_tensor(csd["conv_layer.weight"]))
                for name in ("common", "action", "value"):
                    getattr(pol, name)[0].weight.copy_(torch.as_tensor(csd[f"{name}.0.weight"]))
                    getattr(pol, name)[0].bias.copy_(torch.as_tensor(csd[f"{name}.0.bias"]))
                pol.eval()
                l, v = pol(x) if tag == "plain" else pol(x, perm_indices=torch.as_tensor(perm_idx))
            out[f"dim{conv_dim}.{tag}.logits"] = l.numpy().astype(np.float32)
            out[f"dim{conv_dim}.{tag}.values"] = v.numpy().astype(np.float32).reshape(-1)
        out[f"dim{conv_dim}.seed"] = np.int64(seed)
    np.savez(OUT / "policy_conv1d.npz", **out)

    # ---- deeper layer stacks: BasicPolicy(common_layers=(256,128), policy_layers=(64,), value_layers=(32,))
    out = {}
    rng = np.random.default_rng(152)
    st = scramble_states(rng, 192, 4, 4, 64)
    x = one_hot(st)
    perm_idx = rng.integers(0, 2, size=len(st)).astype(np.int64)
    dsd = synth_deep_state_dict(1520, 256, 512, (256, 128), (64,), (32,), 4)
    out["states"] = st; out["twist_perm_idx"] = perm_idx.astype(np.int32); out["seed"] = np.int64(1520)
    for tag, perms in (("plain", ((), ())), ("twist", (obs_perms, act_perms))):
        pol = BasicPolicy([16, 16], 4, 512, common_layers=(256, 128), policy_layers=(64,), value_layers=(32,),
                          obs_perms=perms[0], act_perms=perms[1], device="cpu")
        pol.load_state_dict({k: torch.as_tensor(v) for k, v in dsd.items()})
        pol.eval()
        with torch.no_grad():
            l, v = pol(x) if tag == "plain" else pol(x, perm_indices=torch.as_tensor(perm_idx))
        out[f"{tag}.logits"] = l.numpy().astype(np.float32)
        out[f"{tag}.values"] = v.numpy().astype(np.float32).reshape(-1)
    np.savez(OUT / "policy_deep.npz", **out)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
