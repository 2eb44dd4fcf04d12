"""Host-side model of the balanced item schedule of k_forward_tc2 (twr_forward_tc2.cu, struct Sched): checks that
every (group, step) is issued exactly once, that a pair never issues the same group twice within two items, and
estimates the makespan (in item times) with the cross-pair hand-offs."""
import sys
MINBASE = int(sys.argv[1]) if len(sys.argv) > 1 else 2   # the kernel requires >= 2 own groups per pair


def make_sched(G, P, T, pair, delta):
    s = dict(T=T, P=P, nA=0, nB=0, sA0=0, pA0=0, gA=0)
    base, rem = G // P, G % P
    if delta > 0 and T >= 4 and base >= MINBASE and rem > 0:
        s["lanes"] = base
        X = rem * T
        Lp = max((X + P - 1) // P, 4)
        x0 = min(X, pair * Lp); x1 = min(X, x0 + Lp); nX = x1 - x0
        if nX > 0:
            s["gA"], s["sA0"] = x0 // T, x0 % T
            s["nA"] = min(nX, T - s["sA0"]); s["nB"] = nX - s["nA"]
            dA = delta * (pair - (s["gA"] * T) // Lp) if s["sA0"] > 0 else 0
            # the piece must end inside the pair's sequence (base * T + nX items): with few own groups the slack granted
            # to late hand-offs is clamped (a shorter slack can only make the consumer wait, never deadlock)
            dA = max(0, min(dA, base * T + nX - 1 - 2 * (s["sA0"] + s["nA"] - 1)))
            s["pA0"] = 2 * s["sA0"] + dA
        s["n_items"] = base * T + nX
    else:
        s["lanes"] = max(0, (G - pair + P - 1) // P)
        s["n_items"] = s["lanes"] * T
    return s


def item(s, pair, i):
    if s["nA"] + s["nB"] == 0:
        return pair + (i % s["lanes"]) * s["P"], i // s["lanes"], 0
    ja = i - s["pA0"]
    if i % 2 == 0 and i // 2 < s["nB"]:
        return s["lanes"] * s["P"] + s["gA"] + 1, i // 2, 2
    if ja >= 0 and ja % 2 == 0 and ja // 2 < s["nA"]:
        return s["lanes"] * s["P"] + s["gA"], s["sA0"] + ja // 2, 1
    before = min(s["nB"], (i + 1) // 2) + (min(s["nA"], (ja + 1) // 2) if ja > 0 else 0)
    m = i - before
    return pair + (m % s["lanes"]) * s["P"], m // s["lanes"], 0


def check(G, P, T, delta, handoff=1.5):
    seen = {}
    scheds = [make_sched(G, P, T, p, delta) for p in range(P)]
    seqs = []
    for p, s in enumerate(scheds):
        seq = [item(s, p, i) for i in range(s["n_items"])]
        seqs.append(seq)
        last = {}
        for i, (g, st, ex) in enumerate(seq):
            assert g < G and 0 <= st < T, (G, P, T, p, i, g, st)
            assert (g, st) not in seen, ("dup", G, P, T, p, i, g, st)
            seen[(g, st)] = (p, i)
            if g in last:
                assert st == last[g][1] + 1 and (i - last[g][0] >= 2 or s["lanes"] == 1), ("spacing", G, P, T, p, i, g)
            last[g] = (i, st)
    assert len(seen) == G * T, (len(seen), G * T)
    # makespan: item i of pair p starts at max(prev end, done(g, st-1) + handoff if produced by another pair)
    done = {}
    t = [0.0] * P
    idx = [0] * P
    progressed = True
    while progressed:
        progressed = False
        for p in range(P):
            while idx[p] < len(seqs[p]):
                g, st, ex = seqs[p][idx[p]]
                start = t[p]
                if st > 0:
                    q, _ = seen[(g, st - 1)]
                    if (g, st - 1) not in done:
                        break
                    if q != p:
                        start = max(start, done[(g, st - 1)] + handoff)
                t[p] = start + 1.0
                done[(g, st)] = t[p]
                idx[p] += 1
                progressed = True
    assert all(idx[p] == len(seqs[p]) for p in range(P)), "deadlock"
    return max(t), max(s["n_items"] for s in scheds)


if __name__ == "__main__":
    P = 74
    for G, T in [(256, 32), (256, 8), (223, 32), (295, 32), (260, 32), (230, 16), (512, 32), (300, 5), (222, 32), (128, 32), (4096, 32)]:
        for delta in (0, 2):
            mk, mx = check(G, P, T, delta)
            print(f"G={G} T={T} delta={delta}: makespan {mk:.1f} items, max load {mx}, ideal {G * T / P:.1f}")
    import random
    random.seed(1)
    for _ in range(3000):
        G = random.randint(1, 700); T = random.randint(1, 40); P = random.choice([2, 7, 66, 74])
        check(G, P, T, 2)
    print("random schedules ok")
