"""Checker used by the GPU parity tests and __graft_entry__.smoke(): replays a device-collected
CollectedData through the CPU oracle (record by record, in the reference's merge order)."""
from __future__ import annotations

import numpy as np

from oracle import orc


def make_policies(sd, obs_size, obs_perms=(), act_perms=()):
    """(device policy, oracle policy) from a BasicPolicy state dict, via the to_rust() layouts."""
    from twisterl_b200 import nn as twn
    pol = twn.Policy(twn.EmbeddingBag(sd["embeddings.weight"].T, sd["embeddings.bias"], True, [obs_size], 0),
                     twn.Sequential([twn.Linear(sd["common.0.weight"].T.flatten(), sd["common.0.bias"], True)]),
                     twn.Sequential([twn.Linear(sd["action.0.weight"].T.flatten(), sd["action.0.bias"], False)]),
                     twn.Sequential([twn.Linear(sd["value.0.weight"].T.flatten(), sd["value.0.bias"], False)]),
                     [list(p) for p in obs_perms], [list(p) for p in act_perms])
    return pol, orc.Policy.from_torch_state_dict(sd, obs_perms, act_perms)


def make_policies_general(sd, obs_size, obs_perms=(), act_perms=()):
    """(device policy, oracle policy) from ANY BasicPolicy state dict (deeper common / head stacks included), the way
    sequential_to_rust exports them (nn/utils.py:17-44): a Linear carries ReLU when a ReLU follows it in the Sequential,
    i.e. every common layer and every head layer but the last."""
    from twisterl_b200 import nn as twn

    def stack(name, relu_last):
        idxs = sorted({int(k.split(".")[1]) for k in sd if k.startswith(name + ".")})
        return twn.Sequential([twn.Linear(sd[f"{name}.{i}.weight"].T.flatten(), sd[f"{name}.{i}.bias"], relu_last or n + 1 < len(idxs))
                               for n, i in enumerate(idxs)])
    pol = twn.Policy(twn.EmbeddingBag(sd["embeddings.weight"].T, sd["embeddings.bias"], True, [obs_size], 0),
                     stack("common", True), stack("action", False), stack("value", False),
                     [list(p) for p in obs_perms], [list(p) for p in act_perms])
    return pol, orc.Policy.from_torch_state_dict(sd, obs_perms, act_perms)


def make_conv1d_policies(sd, obs_shape, conv_dim, obs_perms=(), act_perms=()):
    """(device policy, oracle policy) of a Conv1dPolicy, via Conv1dPolicy.to_rust()'s layouts (nn/policy.py:259-266)."""
    from twisterl_b200 import nn as twn
    w = np.asarray(sd["conv_layer.weight"], np.float32)[:, :, 0]
    pol = twn.Policy(twn.EmbeddingBag(w.T.tolist(), [0.0] * (w.shape[0] * obs_shape[1 - conv_dim]), True, list(obs_shape), conv_dim),
                     twn.Sequential([twn.Linear(sd["common.0.weight"].T.flatten(), sd["common.0.bias"], True)]),
                     twn.Sequential([twn.Linear(sd["action.0.weight"].T.flatten(), sd["action.0.bias"], False)]),
                     twn.Sequential([twn.Linear(sd["value.0.weight"].T.flatten(), sd["value.0.bias"], False)]),
                     [list(p) for p in obs_perms], [list(p) for p in act_perms])
    return pol, orc.Policy.from_conv1d_state_dict(sd, obs_shape, conv_dim, obs_perms, act_perms)


def check_collect_against_oracle(data, spec, opol, seed, collect_id, gamma, lam, tol, env_id_base=0,
                                 max_episodes=None, near_tie=1e-4, stride=1):
    """Replay check (north_star: 'bit-exact under forced or replayed action sequences').

    For every episode (reference merge order): the oracle env, reset from the same Philox stream and
    stepped with the RECORDED actions, must reproduce obs / rewards / terminal flag bit-exactly; the
    recorded masked logits and values must match the oracle forward within `tol`
    (|d| <= tol * max(1, |ref|)); the recorded action must be the Gumbel-max of the recorded logits under
    the shared uniforms (ties closer than `near_tie` are skipped); advs/rets must match the oracle GAE of
    the recorded rewards/values within 1e-5; the twist index must be the shared-stream pick.
    """
    obs, logits = data.obs_array.astype(np.int64), data.logits_array
    values, rewards = data.values_array, data.rewards_array
    actions, perms = data.actions_array.astype(np.int64), data.perms_array.astype(np.int64)
    advs, rets = data.additional_array("advs"), data.additional_array("rets")
    ep_len = np.asarray(data.ep_len)
    n_ep = len(ep_len)
    order = orc.merge_order(n_ep)
    assert int(ep_len.sum()) == len(obs) == len(values) == len(advs), "record count mismatch"
    env = orc.Env(spec)
    n_perms = orc.lib().orc_policy_n_perms(opol._h)
    off = 0
    worst_l = worst_v = worst_g = 0.0
    skipped = checked = 0
    for slot, ep in enumerate(order):
        n = int(ep_len[ep])
        if (max_episodes is not None and slot >= max_episodes * stride) or slot % stride:   # stride: sample every k-th episode
            off += n
            continue
        env.reset(seed=seed, env_id=env_id_base + int(ep), collect_id=collect_id)
        for t in range(n):
            r = off + t
            assert env.observe() == obs[r].tolist(), f"obs mismatch ep {ep} t {t}"
            assert np.float32(env.reward()) == rewards[r], f"reward mismatch ep {ep} t {t}"
            assert env.is_final() == (t == n - 1), f"terminal flag mismatch ep {ep} t {t}"
            perm = -1
            if n_perms:
                w = orc.philox([env_id_base + int(ep), t, orc.RNG_PERM, collect_id], [seed & 0xFFFFFFFF, seed >> 32])
                perm = (int(w[0]) * n_perms) >> 32
            assert perm == perms[r], f"twist mismatch ep {ep} t {t}"
            ml, v = opol.forward(env.observe(), env.masks(), perm)
            masked = ml == np.float32(-1e10)
            assert np.array_equal(masked, logits[r] == np.float32(-1e10)), f"mask mismatch ep {ep} t {t}"
            dl = np.abs(logits[r] - ml)[~masked] / np.maximum(1.0, np.abs(ml[~masked]))
            worst_l = max(worst_l, float(dl.max(initial=0.0)))
            worst_v = max(worst_v, abs(float(values[r]) - float(v)) / max(1.0, abs(float(v))))
            w = orc.philox([env_id_base + int(ep), t, orc.RNG_SAMPLE, collect_id], [seed & 0xFFFFFFFF, seed >> 32])
            u = np.array([orc.u32_to_unit_f32(int(x)) for x in w], dtype=np.float32)
            with np.errstate(divide="ignore"):
                g = logits[r].astype(np.float64) - np.log(np.abs(np.log(u.astype(np.float64))))
            top = np.sort(g)[::-1]
            if top[0] - top[1] < near_tie:
                skipped += 1
            else:
                assert orc.sample_from_logits(logits[r], u) == actions[r], f"action mismatch ep {ep} t {t}"
            checked += 1
            env.step(int(actions[r]))
        a, g_ = orc.gae(rewards[off:off + n], values[off:off + n], gamma, lam)
        worst_g = max(worst_g, float(np.abs(a - advs[off:off + n]).max()), float(np.abs(g_ - rets[off:off + n]).max()))
        off += n
    assert worst_l <= tol, f"logits off by {worst_l} (tol {tol})"
    assert worst_v <= tol, f"values off by {worst_v} (tol {tol})"
    assert worst_g <= 1e-5, f"GAE off by {worst_g}"
    return dict(records=checked, near_ties_skipped=skipped, max_logit_err=worst_l, max_value_err=worst_v,
                max_gae_err=worst_g)
