"""GPU parity of collector.evaluate / collector.solve against the oracle's
restatement of rust/src/rl/solve.rs and rust/src/rl/evaluate.rs on the shared Philox streams."""
import ctypes as C
import os

import numpy as np
import pytest

from suite_loader import suite_precision

from helpers import synth_state_dict, trained15, transpose_twists
from oracle import orc

pytestmark = pytest.mark.gpu
PRECISION = suite_precision(globals())


@pytest.fixture(scope="module")
def eng():
    import twisterl_b200 as tw
    tw.configure(device=0, precision=PRECISION, seed=0x1234ABCD)
    return tw.default_engine()


# Device and oracle evaluate the same f32 expressions on the same Philox streams; their logits differ in the last bits, so
# an episode may differ only where one of its decisions was a near-tie in the oracle (argmax gap, weighted-draw distance
# to a bin edge, UCB gap inside MCTS; orc_evaluate_margins reports the smallest of an episode).
NEAR_TIE = {"fp32": 2e-4, "f16x2": 2e-3, "f16x2w16": 2e-2, "f16f8c": 2e-2}[PRECISION]


def _evaluate(eng, env, pol, n, det, searches, cid, mcts=0, c_puct=1.41, depth=1, episodes=False):
    from twisterl_b200 import _lib
    import twisterl_b200 as tw
    spec = tw.env.spec_from_env(env)
    s, r = C.c_float(), C.c_float()
    bs, bt = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32)
    eng.set_collect_id(cid)
    _lib.check(_lib.load().twr_evaluate_episodes(eng._h, C.byref(spec), pol.device_handle(eng), n, int(det), searches, mcts, c_puct,
                                                 depth, C.byref(s), C.byref(r), _lib.ptr(bs), _lib.ptr(bt)))
    if episodes:
        return float(s.value), float(r.value), bs[:n], bt[:n]
    return float(s.value), float(r.value)


def _check_episodes(bs, bt, obs, obt, margins, what):
    """per-episode best (success, reward): equal to the oracle's, or the oracle met a near-tie in that episode"""
    diff = [i for i in range(len(bs)) if bs[i] != obs[i] or abs(float(bt[i]) - float(obt[i])) > 1e-5]
    for i in diff:
        assert margins[i] < NEAR_TIE, f"{what}: episode {i} differs ({bs[i]}, {bt[i]}) vs ({obs[i]}, {obt[i]}) with smallest margin {margins[i]}"
    print(f"[{what}] {len(bs) - len(diff)}/{len(bs)} episodes identical, {len(diff)} explained by near-ties")
    assert len(diff) <= len(bs) // 4


@pytest.mark.parametrize("det,searches", [(True, 1), (False, 1), (False, 6)])
def test_evaluate_matches_oracle(eng, det, searches):
    import twisterl_b200 as tw
    from parity import make_policies
    _, sd = trained15()
    pol, opol = make_policies(sd, 256)
    ospec = orc.puzzle_spec(4, 4, 10, 2, 256)
    env = tw.env.Puzzle(4, 4, 10, 2, 256)
    n = 200
    s, r, bs, bt = _evaluate(eng, env, pol, n, det, searches, cid=7, episodes=True)
    obs_, obt, mm = orc.evaluate_margins(ospec, opol, n, det, searches, seed=eng.seed, collect_id=7)
    _check_episodes(bs, bt, obs_, obt, mm, f"evaluate det={det} searches={searches}")
    assert abs(s - float(bs.mean())) < 1e-6 and abs(r - float(bt.mean())) < 1e-4          # the means are the means of these
    assert 0.0 <= s <= 1.0


def test_evaluate_edge_cases_and_api(eng):
    import twisterl_b200 as tw
    from parity import make_policies
    _, sd = trained15()
    pol, _ = make_policies(sd, 256)
    env = tw.env.Puzzle(4, 4, 0, 2, 256)                      # difficulty 0: already solved, reward 1.0
    assert tw.collector.evaluate(env, pol, 16, False, 1, 0, 0, 1.4, 1, 32) == (1.0, 1.0)
    s, r = _evaluate(eng, env, pol, 8, False, 0, cid=1)       # zero searches: solve()'s initial best
    assert s == 0.0 and r == float("-inf")
    assert tw.collector.evaluate(env, pol, 4, False, 1, 3, 0, 1.4, 1, 1) == (1.0, 1.0)     # MCTS-guided, already solved
    # keyword form used by rl/algorithm.py:98
    out = tw.collector.evaluate(env, pol, num_episodes=4, deterministic=True, num_searches=2, num_mcts_searches=0, seed=1,
                                C=1.4, max_expand_depth=1, num_cores=4)
    assert out == (1.0, 1.0)
    g = tw.env.GridWorld(5, 5, 64, 6)
    gp, _ = make_policies(synth_state_dict(5, 625, 512, 128, 4), 625)
    s, r = tw.collector.evaluate(g, gp, 64, False, 2, 0, 0, 1.4, 1, 1)    # 625-row table: fp32 kernel on either engine
    assert 0.0 <= s <= 1.0 and -40.0 < r <= 1.0


def test_solve_matches_oracle_and_replays(eng):
    import twisterl_b200 as tw
    from parity import make_policies
    _, sd = trained15()
    obs_perms, act_perms = transpose_twists(4)
    for perms in (((), ()), (obs_perms, act_perms)):
        pol, opol = make_policies(sd, 256, *perms)
        env = tw.env.Puzzle(4, 4, 1, 2, 256)
        start = [1, 5, 2, 3, 4, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]     # two moves from solved
        env.set_state(start)
        (succ, rew), acts = tw.collector.solve(env, pol, True, 1, 0, 1.4, 1)
        assert env.get_state() == start                                   # the env itself is not advanced
        oenv = orc.Env(orc.puzzle_spec(4, 4, 1, 2, 256)); oenv.set_state(start)
        if not perms[0]:
            (os_, or_), oacts = orc.solve(oenv, opol, True, 1, seed=eng.seed)
            assert (succ, acts) == (os_, oacts) and abs(rew - or_) < 1e-6
        assert succ == 1.0 and len(acts) >= 2
        # the returned action list replays to a solved board with the returned reward
        tot = np.float32(0)
        for a in acts:
            tot = np.float32(tot + np.float32(oenv.reward()))
            oenv.step(a)
        tot = np.float32(tot + np.float32(oenv.reward()))
        assert oenv.success() and abs(float(tot) - rew) < 1e-6
    # the 8-puzzle of the reference notebook (examples/puzzle.ipynb): sampled best-of-100 from a hard start
    pol8, _ = make_policies(synth_state_dict(8, 81, 512, 256, 4), 81)
    env8 = tw.env.Puzzle(3, 3, 1, 2, 256)
    env8.set_state([8, 7, 5, 3, 2, 0, 4, 6, 1])
    (s8, r8), a8 = tw.collector.solve(env8, pol8, False, 100, 0, 1.4, 1)
    assert s8 in (0.0, 1.0) and len(a8) <= 256
    if s8 == 0.0:
        assert len(a8) == 256 and abs(r8 - (255 * (-0.5 / 256) + (-0.5 / 256) + -0.5)) < 1e-3


def test_collect_torch_device_handoff(eng):
    """f2: torch CUDA tensors aliasing the engine's output must equal the host copy of the same collect."""
    torch = pytest.importorskip("torch")
    import twisterl_b200 as tw
    from parity import make_policies
    _, sd = trained15()
    pol, _ = make_policies(sd, 256)
    env = tw.env.Puzzle(4, 4, 8, 2, 256)
    col = tw.collector.PPOCollector(300, 0.995, 0.995, 1, engine=eng)
    eng.set_collect_id(4)
    ref = col.collect(env, pol)
    eng.set_collect_id(4)
    t = col.collect_torch(env, pol)
    assert t["obs"].is_cuda and t["obs"].shape == (len(ref.values_array), 256)
    dense = np.zeros((len(ref.values_array), 256), np.float32)
    np.put_along_axis(dense, ref.obs_array.astype(np.int64), 1.0, axis=1)
    assert np.array_equal(t["obs"].cpu().numpy(), dense)
    assert np.array_equal(t["logits"].cpu().numpy(), ref.logits_array)
    assert np.array_equal(t["actions"].cpu().numpy(), ref.actions_array.astype(np.int64))
    assert np.array_equal(t["advs"].cpu().numpy(), ref.additional_array("advs"))
    assert np.array_equal(t["rets"].cpu().numpy(), ref.additional_array("rets"))
    assert np.array_equal(t["perms"].cpu().numpy(), ref.perms_array.astype(np.int64))
    assert t["stats"]["records"] == len(ref.values_array)


@pytest.mark.parametrize("det,searches,sims", [(True, 1, 24), (False, 2, 10)])
def test_evaluate_with_mcts_matches_oracle(eng, det, searches, sims):
    """defaults.py's "mcts_100" evaluation: single_solve takes its action distribution from predict_probs_mcts
    (rl/solve.rs:37-48).  Same Philox streams on both sides; a flipped UCB near-tie can change single rollouts."""
    import twisterl_b200 as tw
    from parity import make_policies
    pol, opol = make_policies(synth_state_dict(8, 81, 512, 256, 4), 81)
    ospec = orc.puzzle_spec(3, 3, 4, 2, 256)
    env = tw.env.Puzzle(3, 3, 4, 2, 256)
    n = 96
    s, r, bs, bt = _evaluate(eng, env, pol, n, det, searches, cid=13, mcts=sims, c_puct=1.41, depth=1, episodes=True)
    obs_, obt, mm = orc.evaluate_margins(ospec, opol, n, det, searches, seed=eng.seed, collect_id=13, num_mcts_searches=sims,
                                         c_puct=1.41, max_expand_depth=1)
    _check_episodes(bs, bt, obs_, obt, mm, f"evaluate+mcts det={det} searches={searches} sims={sims}")
    # the search must help: MCTS-guided evaluation solves at least as often as the raw synthetic policy
    s0, _ = _evaluate(eng, env, pol, n, det, searches, cid=13)
    assert s >= s0 - 2.0 / n


def test_solve_with_mcts(eng):
    import twisterl_b200 as tw
    from parity import make_policies
    _, sd = trained15()
    pol, opol = make_policies(sd, 256)
    env = tw.env.Puzzle(4, 4, 1, 2, 256)
    start = [1, 5, 2, 3, 4, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]         # two moves from solved
    env.set_state(start)
    (succ, rew), acts = tw.collector.solve(env, pol, True, 1, 16, 1.41, 1)
    oenv = orc.Env(orc.puzzle_spec(4, 4, 1, 2, 256)); oenv.set_state(start)
    (os_, or_), oacts = orc.solve(oenv, opol, True, 1, seed=eng.seed, num_mcts_searches=16, c_puct=1.41, max_expand_depth=1)
    assert succ == 1.0 and os_ == 1.0
    assert acts == oacts and abs(rew - or_) < 1e-6
    for a in acts:
        oenv.step(a)
    assert oenv.success()


def test_engine_and_policy_lifecycle_does_not_leak():
    """Engines, policies, env batches and collects created and destroyed repeatedly give their device memory back
    (every cudaMalloc of the library has an owner that frees it)."""
    torch = pytest.importorskip("torch")
    import twisterl_b200 as tw
    from parity import make_policies
    _, sd = trained15()

    def cycle():
        eng = tw.Engine(device=0, precision=PRECISION, seed=5)
        pol, _ = make_policies(sd, 256)
        env = tw.env.Puzzle(4, 4, 4, 2, 256)
        tw.collector.PPOCollector(3000, 0.995, 0.995, 1, engine=eng).collect(env, pol)
        pol8, _ = make_policies(synth_state_dict(8, 81, 512, 256, 4), 81)
        tw.collector.AZCollector(64, 8, 1.41, 1, 1, engine=eng).collect(tw.env.Puzzle(3, 3, 3, 2, 256), pol8)
        pol.release(); pol8.release()
        eng.close()

    cycle()                                   # warm: CUDA context, module load
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(6):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 128 << 20, f"device memory shrank by {(free0 - free1) >> 20} MiB over 6 engine lifecycles"
