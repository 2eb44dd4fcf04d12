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


def _mcts(eng, pol, batch, n_sims, c, med, base, cid, t):
    from twisterl_b200 import _lib
    n, A = batch.n, 4
    probs = np.zeros((n, A), np.float32); visits = np.zeros((n, A), np.int32)
    _lib.check(_lib.load().twr_mcts_probs(eng._h, pol.device_handle(eng), batch._h, n_sims, c, med, base, cid, t,
                                          _lib.ptr(probs), _lib.ptr(visits)))
    return probs, visits


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
    probs, visits = _mcts(eng, pol, b, n_sims, 1.41, med, 500, 3, 2)
    same = 0
    for i in range(n):
        env = orc.Env(orc.puzzle_spec(4, 4, 5, 2, 256)); env.set_state(states[i])
        op, ov = orc.mcts_probs(env, opol, n_sims, 1.41, med, seed=eng.seed, collect_id=3, stream_id=500 + i, t=2)
        assert visits[i].sum() == ov.sum() == (n_sims if not env.is_final() or n_sims == 0 or True else 0) or env.is_final()
        assert abs(probs[i].sum() - 1.0) < 1e-5
        masks = env.masks()
        if not env.is_final():
            assert all(visits[i][a] == 0 for a in range(4) if not masks[a])
        same += int(np.array_equal(visits[i], ov))
        # a flipped UCB near-tie (ulp-level exp/forward differences) can move a few visits, never the bulk
        assert np.abs(probs[i] - op).sum() <= 0.35
    assert same >= 0.9 * n, same


def test_az_collect_structure_and_replay(eng):
    import twisterl_b200 as tw
    from parity import make_policies
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
            assert o.observe() == d.obs_array[off + t].tolist()
            assert np.float32(o.reward()) == d.step_rewards[off + t]
            assert o.is_final() == (t == n - 1)
            masks = o.masks()
            assert all(probs[off + t][a] == 0 for a in range(4) if not masks[a])
            assert probs[off + t][int(d.step_actions[off + t])] > 0          # the drawn action has MCTS mass
            op, _ = orc.mcts_probs(o, opol, sims, 1.41, 1, seed=eng.seed, collect_id=9, stream_id=int(ep), t=t)
            match += int(np.allclose(op, probs[off + t], atol=1e-6))
            prefix.append(tot); tot = np.float32(tot + d.step_rewards[off + t])
            o.step(int(d.step_actions[off + t]))
        want = np.array([np.float32(tot - q) for q in prefix], dtype=np.float32)
        assert np.array_equal(want, d.additional_array("remaining_values")[off:off + n])
        off += n
    assert match >= 0.9 * R
    # whole-collect comparison with the oracle collector on the same streams
    oc = orc.az_collect(ospec, opol, E, sims, 1.41, 1, seed=eng.seed, collect_id=9)
    assert sum(int(a == b) for a, b in zip(oc["ep_len"], d.ep_len)) >= 0.8 * E


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
