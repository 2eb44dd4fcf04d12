"""GPU parity of the AlphaZero path: batched MCTS and AZCollector against the
oracle's restatement of rust/src/rl/search.rs and rust/src/collector/az.rs on the shared Philox streams."""
import ctypes as C
import os

import numpy as np
import pytest

from suite_loader import suite_precision

from helpers import scramble_states, synth_state_dict, trained15
from oracle import orc

pytestmark = pytest.mark.gpu
PRECISION = suite_precision(globals())


@pytest.fixture(scope="module")
def eng():
    import twisterl_b200 as tw
    e = tw.Engine(device=0, precision=PRECISION, seed=0xA11CE)
    yield e
    e.close()


# A device search and the oracle's run the same f32 expressions on the same Philox streams; their leaf evaluations differ
# in the last bits (1e-6-grade on the fp32 path, 1e-5-grade logits on the split-fp16 tensor-core path), which can flip a
# decision only where the oracle saw a near-tie.  Every search that does not reproduce the oracle's visit counts exactly
# is therefore traced to the FIRST simulation in which the two part ways, and the oracle's smallest decision margin of
# that simulation (best minus second-best UCB along the descent, bin-edge distance of the child draw) must be tiny.
# UCB = q + C*sqrt(N)/(1+n)*prior amplifies a prior difference by up to C*sqrt(N) ~ 20 at these search sizes.
NEAR_TIE = {"fp32": 2e-4, "f16x2": 2e-3, "f16x2w16": 2e-2, "f16f8c": 2e-2}[PRECISION]


def _mcts(eng, pol, batch, n_sims, c, med, base, cid, t):
    """(probs [n][A], visits [n][A], trace [n_sims][n][2]) of twr_debug_mcts_trace."""
    from twisterl_b200 import _lib
    n, A = batch.n, 4
    probs = np.zeros((n, A), np.float32); visits = np.zeros((n, A), np.int32)
    trace = np.zeros((max(n_sims, 1), n, 2), np.int32)
    _lib.check(_lib.load().twr_debug_mcts_trace(eng._h, pol.device_handle(eng), batch._h, n_sims, c, med, base, cid, t,
                                                _lib.ptr(probs), _lib.ptr(visits), _lib.ptr(trace)))
    return probs, visits, trace[:n_sims]


def _first_divergence_margin(dev_trace, oenv, opol, n_sims, c, med, seed, cid, stream_id, t):
    """Oracle margin of the first simulation whose (leaf, backed-up node) differs from the device's; None if none does."""
    _, _, leaf, child, margin = orc.mcts_trace(oenv, opol, n_sims, c, med, seed=seed, collect_id=cid, stream_id=stream_id, t=t)
    for s in range(n_sims):
        if dev_trace[s][0] != leaf[s] or dev_trace[s][1] != child[s]:
            return s, float(margin[s])
    return None


@pytest.mark.parametrize("n_sims,med", [(0, 1), (24, 1), (40, 2), (16, 0), (200, 1)])
def test_mcts_probs_match_oracle(eng, n_sims, med):
    from parity import make_policies
    from twisterl_b200 import _lib
    from twisterl_b200.env import EnvBatch
    _, sd = trained15()
    pol, opol = make_policies(sd, 256)
    rng = np.random.default_rng(n_sims + med)
    n = 96
    states = scramble_states(rng, n, 4, 4, 12)
    b = EnvBatch(_lib.EnvSpec(0, 4, 4, 5, 2, 256), n, eng)
    b.set_state(states)
    probs, visits, trace = _mcts(eng, pol, b, n_sims, 1.41, med, 500, 3, 2)
    same, worst = 0, 0.0
    for i in range(n):
        env = orc.Env(orc.puzzle_spec(4, 4, 5, 2, 256)); env.set_state(states[i])
        op, ov = orc.mcts_probs(env, opol, n_sims, 1.41, med, seed=eng.seed, collect_id=3, stream_id=500 + i, t=2)
        assert visits[i].sum() == ov.sum()                       # one backup per simulation on both sides
        assert abs(probs[i].sum() - 1.0) < 1e-5
        masks = env.masks()
        if not env.is_final():
            assert all(visits[i][a] == 0 for a in range(4) if not masks[a])
        if np.array_equal(visits[i], ov):
            same += 1
            assert np.allclose(probs[i], op, atol=1e-6)
            continue
        div = _first_divergence_margin(trace[:, i], env, opol, n_sims, 1.41, med, eng.seed, 3, 500 + i, 2)
        assert div is not None, f"env {i}: visit counts differ but every simulation took the oracle's path"
        assert div[1] < NEAR_TIE, f"env {i}: first divergence at simulation {div[0]} where the oracle's margin was {div[1]}"
        worst = max(worst, div[1])
    print(f"[mcts n_sims={n_sims} med={med}] identical {same}/{n}, largest margin at a divergence {worst:.2e}")
    assert same >= n // 2                                         # near-ties are the exception, not the rule


def _weighted_index_f32(w, u):
    """rand's WeightedIndex as nn/policy.rs:153-167 uses it, in f32 like the device and the oracle."""
    w = np.asarray(w, np.float32)
    tw = np.float32(0)
    for x in w:
        tw = np.float32(tw + x)
    if not tw > 0:
        return 0
    chosen = np.float32(np.float32(u) * tw)
    cum, last = np.float32(0), 0
    for i, x in enumerate(w):
        if x > 0:
            cum = np.float32(cum + x); last = i
            if cum > chosen:
                return i
    return last


def test_az_collect_structure_and_replay(eng):
    """Every record of a device AZ collect, replayed through the oracle in merge order: observation, reward and terminal
    flag bit-exact; the stored MCTS distribution equal to the oracle's or explained by a near-tie at the first divergent
    simulation; the recorded action the exact WeightedIndex draw of the stored distribution under the shared uniform;
    remaining_values the f32 reward-to-go of az.rs:93."""
    import twisterl_b200 as tw
    from parity import make_policies
    from twisterl_b200 import _lib
    from twisterl_b200.env import EnvBatch
    sd = synth_state_dict(4, 81, 512, 128, 4)
    pol, opol = make_policies(sd, 81)
    ospec = orc.puzzle_spec(3, 3, 3, 2, 256)
    env = tw.env.Puzzle(3, 3, 3, 2, 256)
    E, sims = 40, 12
    col = tw.collector.AZCollector(E, sims, 1.41, 1, 32, engine=eng)
    eng.set_collect_id(9)
    d = col.collect(env, pol)
    R = len(d.obs_array)
    assert R == int(d.ep_len.sum()) and d.values == [] and d.rewards == [] and d.actions == []
    assert set(d.additional_data) == {"remaining_values"} and d.perms == [-1] * R
    probs = d.logits_array
    assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-5)
    order = orc.merge_order(E)
    off = 0
    match = 0
    for ep in order:
        n = int(d.ep_len[ep])
        o = orc.Env(ospec); o.reset(seed=eng.seed, env_id=int(ep), collect_id=9)
        tot = np.float32(0); prefix = []
        for t in range(n):
            r = off + t
            assert o.observe() == d.obs_array[r].tolist()
            assert np.float32(o.reward()) == d.step_rewards[r]
            assert o.is_final() == (t == n - 1)
            masks = o.masks()
            assert all(probs[r][a] == 0 for a in range(4) if not masks[a])
            op, _ = orc.mcts_probs(o, opol, sims, 1.41, 1, seed=eng.seed, collect_id=9, stream_id=int(ep), t=t)
            if np.allclose(op, probs[r], atol=1e-6):
                match += 1
            else:
                # rebuild this state on the device (same reset stream, the recorded actions) and trace the search
                b = EnvBatch(_lib.EnvSpec(0, 3, 3, 3, 2, 256), 1, eng)
                b.reset(env_id_base=int(ep), collect_id=9)
                for a in d.step_actions[off:r]:
                    b.step([int(a)])
                assert b.get_state()[0].tolist() == o.get_state()
                p2, _, trace = _mcts(eng, pol, b, sims, 1.41, 1, int(ep), 9, t)
                assert np.array_equal(p2[0], probs[r])           # the parity API reruns the collect's search exactly
                div = _first_divergence_margin(trace[:, 0], o, opol, sims, 1.41, 1, eng.seed, 9, int(ep), t)
                assert div is not None and div[1] < NEAR_TIE, (int(ep), t, div)
            w = orc.philox([int(ep), t, 5, 9], [eng.seed & 0xFFFFFFFF, eng.seed >> 32])      # TWR_RNG_AZ_ACT = 5
            assert _weighted_index_f32(probs[r], orc.u32_to_unit_f32(int(w[0]))) == int(d.step_actions[r])   # az.rs:72
            prefix.append(tot); tot = np.float32(tot + d.step_rewards[r])
            o.step(int(d.step_actions[r]))
        want = np.array([np.float32(tot - q) for q in prefix], dtype=np.float32)
        assert np.array_equal(want, d.additional_array("remaining_values")[off:off + n])
        off += n
    print(f"[az collect] {match}/{R} records with the oracle's distribution, the rest explained by near-ties")
    assert match >= R // 2


def test_az_errors_and_trained_policy_solves(eng):
    import twisterl_b200 as tw
    from helpers import transpose_twists
    from parity import make_policies
    _, sd = trained15()
    pol, _ = make_policies(sd, 256)
    env = tw.env.Puzzle(4, 4, 4, 2, 256)
    col = tw.collector.AZCollector(64, 32, 1.41, 1, 1, engine=eng)
    d = col.collect(env, pol)
    assert d.stats["successes"] >= 60                         # MCTS + trained policy solves shallow scrambles
    twisted, _ = make_policies(sd, 256, *transpose_twists(4))
    with pytest.raises(RuntimeError, match="twists"):
        col.collect(env, twisted)
    with pytest.raises(RuntimeError, match="No data in collected data chunks to merge"):
        tw.collector.AZCollector(0, 4, 1.0, 1, 1, engine=eng).collect(env, pol)


@pytest.mark.parametrize("lockstep", [False, True])
def test_mcts_probs_gridworld(eng, monkeypatch, lockstep):
    """The search on the other env kind (GridWorld 5x5: 625-row table, common width 128), on both device paths: the
    persistent whole-search kernel (policy stationary in shared memory) and the lockstep kernels."""
    from parity import make_policies
    from twisterl_b200 import _lib
    from twisterl_b200.env import EnvBatch
    if lockstep:
        monkeypatch.setenv("TWISTERL_B200_MCTS_LOCKSTEP", "1")
    pol, opol = make_policies(synth_state_dict(6, 625, 512, 128, 4), 625)
    n, sims = 64, 60
    b = EnvBatch(_lib.EnvSpec(1, 5, 5, 6, 0, 64), n, eng)
    b.reset(env_id_base=900, collect_id=4)
    states = b.get_state()
    probs, visits, trace = _mcts(eng, pol, b, sims, 1.41, 1, 900, 4, 0)
    same = 0
    for i in range(n):
        env = orc.Env(orc.gridworld_spec(5, 5, 64, 6)); env.reset(seed=eng.seed, env_id=900 + i, collect_id=4)
        assert env.get_state() == states[i].tolist()
        op, ov = orc.mcts_probs(env, opol, sims, 1.41, 1, seed=eng.seed, collect_id=4, stream_id=900 + i, t=0)
        assert visits[i].sum() == ov.sum()
        if np.array_equal(visits[i], ov):
            same += 1
            continue
        div = _first_divergence_margin(trace[:, i], env, opol, sims, 1.41, 1, eng.seed, 4, 900 + i, 0)
        assert div is not None and div[1] < NEAR_TIE, (i, div)
    assert same >= n // 2
