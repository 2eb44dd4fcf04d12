"""End-to-end check that the engine's rollouts train a policy: a compact PyTorch PPO loop (the reference's loss,
src/twisterl/rl/ppo.py:63-113, and curriculum, rl/algorithm.py:165-171) on the 8-puzzle, with collection and
evaluation on the B200 engine: the collect is handed to torch on the device (`collect_torch`) and the weights go back
device-to-device (`Policy.update_from_torch`).  Not part of the product path; the reference's own trainer runs unmodified on
`twisterl_b200.install_as_twisterl()`.

    python examples/train_ppo_puzzle8.py [iterations]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import twisterl_b200 as tw
from twisterl_b200 import nn as twn


class TorchPolicy(torch.nn.Module):          # same architecture as the reference's BasicPolicy (nn/policy.py:25-58)
    def __init__(self, obs_size=81, emb=512, hidden=256, n_act=4):
        super().__init__()
        self.embeddings = torch.nn.Linear(obs_size, emb)
        self.common = torch.nn.Sequential(torch.nn.Linear(emb, hidden), torch.nn.ReLU())
        self.action = torch.nn.Sequential(torch.nn.Linear(hidden, n_act))
        self.value = torch.nn.Sequential(torch.nn.Linear(hidden, 1))

    def forward(self, x):
        h = self.common(torch.relu(self.embeddings(x)))
        return self.action(h), self.value(h)

    def to_engine(self):                     # the to_rust() hand-off (nn/utils.py:17-59)
        g = lambda t: t.detach().cpu().numpy()
        return twn.Policy(twn.EmbeddingBag(g(self.embeddings.weight).T, g(self.embeddings.bias), True, [81], 0),
                          twn.Sequential([twn.Linear(g(self.common[0].weight).T.flatten(), g(self.common[0].bias), True)]),
                          twn.Sequential([twn.Linear(g(self.action[0].weight).T.flatten(), g(self.action[0].bias), False)]),
                          twn.Sequential([twn.Linear(g(self.value[0].weight).T.flatten(), g(self.value[0].bias), False)]), [], [])


def main(iters=60):
    torch.manual_seed(0)
    dev = torch.device("cuda")
    tw.configure(device=0, precision="f16x2", seed=1)
    policy = TorchPolicy().to(dev)
    opt = torch.optim.Adam(policy.parameters(), lr=1.5e-4)
    env = tw.env.Puzzle(3, 3, 1, 2, 256)
    collector = tw.collector.PPOCollector(num_episodes=4096, gamma=0.995, num_cores=32, **{"lambda": 0.995})
    rs = policy.to_engine()                  # built once; refreshed in place from the live CUDA parameters afterwards
    t0 = time.time()
    for it in range(iters):
        rs.update_from_torch(policy)         # f3: device-to-device weight sync (no to_rust() round trip)
        succ, rew = tw.collector.evaluate(env, rs, 256, False, 1, 0, 0, 1.41, 1, 32)
        data = collector.collect_torch(env, rs)      # f2: torch CUDA tensors aliasing the engine's output
        obs, logits, acts, rets, advs = data["obs"], data["logits"], data["actions"], data["rets"], data["advs"]
        advs = (advs - advs.mean()) / (advs.std() + 1e-8)
        old_lp = torch.distributions.Categorical(logits=logits).log_prob(acts)
        illegal = logits <= -1e9
        for _ in range(10):
            pl, pv = policy(obs)
            dist = torch.distributions.Categorical(logits=pl.masked_fill(illegal, -1e10))
            ratio = torch.exp(dist.log_prob(acts) - old_lp)
            loss = (-torch.min(ratio * advs, torch.clamp(ratio, 0.9, 1.1) * advs).mean()
                    + 0.8 * torch.nn.functional.mse_loss(pv.squeeze(1), rets) - 0.01 * dist.entropy().mean())
            opt.zero_grad(); loss.backward(); opt.step()
        print(f"it {it:3d} difficulty {env.difficulty:2d} success {succ:.2f} reward {rew:+.3f} records {len(acts)} "
              f"loss {loss.item():+.4f} ({time.time() - t0:.1f}s)", flush=True)
        if succ >= 0.85 and env.difficulty < 32:
            env.difficulty = env.difficulty + 1          # curriculum, rl/algorithm.py:165-171
    return env.difficulty


if __name__ == "__main__":
    d = main(int(sys.argv[1]) if len(sys.argv) > 1 else 60)
    print("final difficulty", d)
