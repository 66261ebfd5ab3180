"""Model check of the row-pipelined NHWC kernel's hand-off protocol (csrc/ts_nhwc.cu, k_gather_nhwc_rows): a producer that
fills a ring of K row slots in order, W consumer warps that take (output row, pass) tasks round-robin, `full` / `empty`
mbarriers per slot with PARITY waits.  The model restates the kernel's window / segment formulas and replays the W + 1
state machines under randomised and adversarial schedules.  It pins two properties the GPU cannot show by passing tests:
no deadlock, and no warp ever waits on a barrier that is two phases ahead of it (the parity test would then block for ever --
the hang of the first version of the kernel, which released rows before it had seen them land)."""
import random

import pytest


def clampi(v, lo, hi):
    return lo if v < lo else (hi if v > hi else v)


def segments(r0, r1, os0, s0, lb0, smin, smax, win):
    """rows_segment(): the pieces of a CTA's range of global output rows, one per image, with the input rows they reach"""
    out, r = [], r0
    while r < r1:
        n = r // os0
        o_a = r - n * os0
        o_b = min(os0, o_a + (r1 - r))
        r += o_b - o_a
        top, reach = o_b - 1 + lb0 - smax + win - 1, o_b - 1 + lb0 - smin
        lo = clampi(o_a + lb0 - smax, 0, s0 - 1)
        hi = max(lo, clampi(min(top, reach), 0, s0 - 1))
        out.append((o_a, o_b, lo, hi))
    return out


def window(sg, o, lb0, smin, smax, win):
    """rows_window(): the input rows an output row may read from the ring"""
    _, _, lo, hi = sg
    a = o + lb0 - smax
    wlo = clampi(a, lo, hi)
    return wlo, max(wlo, clampi(min(a + win - 1, o + lb0 - smin), lo, hi))


def replay(r0, r1, os0, s0, lb0, smin, smax, K, W, P, seed, observe_first=True, producer_first=False):
    win = K - (W + P - 1) // P - 1
    segs = segments(r0, r1, os0, s0, lb0, smin, smax, win)
    total = sum(hi - lo + 1 for _, _, lo, hi in segs)
    tasks, qbase = [[] for _ in range(W)], 0
    for sg in segs:
        for w in range(W):
            for t in range(w, (sg[1] - sg[0]) * P, W):
                wlo, whi = window(sg, sg[0] + t // P, lb0, smin, smax, win)
                tasks[w].append((qbase + wlo - sg[2], qbase + whi - sg[2]))
        qbase += sg[3] - sg[2] + 1
    fills, frees, arrivals, prod_q = [0] * K, [0] * K, [0] * K, 0
    st = [dict(i=0, rel=0, seen=0) for _ in range(W)]
    rnd = random.Random(seed)

    def observe(s):                      # one parity wait on full[slot]; False = would block, raises on a two-phase gap
        slot, phase = s['seen'] % K, s['seen'] // K
        if fills[slot] <= phase:
            return False
        assert fills[slot] - phase < 2, "parity wait on a barrier two phases ahead: blocks for ever"
        s['seen'] += 1
        return True

    def release(s):
        slot = s['rel'] % K
        arrivals[slot] += 1
        if arrivals[slot] == W:
            arrivals[slot] = 0
            frees[slot] += 1
        s['rel'] += 1

    while True:
        progressed = False
        order = list(range(W + 1))
        rnd.shuffle(order)
        if producer_first:
            order.remove(W)
            order.insert(0, W)
        for a in order:
            if a == W:                   # producer: slot q % K is free once round q // K - 1 has been released by all warps
                while prod_q < total and (prod_q < K or frees[prod_q % K] >= prod_q // K):
                    fills[prod_q % K] += 1
                    prod_q += 1
                    progressed = True
                    if not producer_first:
                        break
                continue
            s = st[a]
            goal_rel, goal_seen = (tasks[a][s['i']][0], tasks[a][s['i']][1] + 1) if s['i'] < len(tasks[a]) else (total, total)
            if s['rel'] >= total and s['i'] >= len(tasks[a]):
                continue
            blocked = False
            while s['rel'] < goal_rel:   # rows below the window: (see them land, then) release them
                if observe_first and s['seen'] <= s['rel'] and not observe(s):
                    blocked = True
                    break
                release(s)
                progressed = True
                if not observe_first and producer_first:
                    break                # adversary: let the producer run between a release and the next look at a barrier
            if blocked or s['rel'] < goal_rel:
                continue
            while s['seen'] < goal_seen:
                if not observe(s):
                    blocked = True
                    break
                progressed = True
            if not blocked and s['i'] < len(tasks[a]):
                s['i'] += 1              # the gather of this task
                progressed = True
        if prod_q == total and all(s['rel'] >= total and s['i'] >= len(tasks[w]) for w, s in enumerate(st)):
            return True
        assert progressed, "deadlock"


CASES = [  # (output rows per image, input rows, lb0, smin, smax, K, W, P)
    (56, 56, 0, -3, 3, 15, 12, 2), (56, 56, 0, -1, 1, 15, 12, 2), (56, 56, 0, -6, 6, 15, 12, 2), (56, 56, 0, -3, 3, 15, 6, 1),
    (9, 9, 0, -5, 5, 9, 4, 1), (20, 24, 3, -5, 5, 9, 4, 1), (1, 1, 0, -2, 2, 7, 2, 1), (5, 12, 4, -2, 2, 7, 8, 4), (56, 56, 0, 0, 0, 64, 12, 2)]


@pytest.mark.parametrize("case", CASES)
def test_observe_then_release_never_blocks(case):
    os0, s0, lb0, smin, smax, K, W, P = case
    rows = 256 * os0
    for b in range(0, 148, 9):
        r0, r1 = rows * b // 148, rows * (b + 1) // 148
        for seed in range(3):
            assert replay(r0, r1, os0, s0, lb0, smin, smax, K, W, P, seed)
        assert replay(r0, r1, os0, s0, lb0, smin, smax, K, W, P, 0, producer_first=True)


def test_release_before_observe_can_block_for_ever():
    """The first version released the rows a warp's windows skip before looking at their barriers.  With a fast producer the
    freed slot is refilled, the refill lands, and the late parity wait faces a barrier two phases ahead."""
    with pytest.raises(AssertionError, match="two phases ahead"):
        for b in range(0, 148, 3):
            rows = 256 * 56
            replay(rows * b // 148, rows * (b + 1) // 148, 56, 56, 0, -1, 1, 15, 8, 1, 0, observe_first=False, producer_first=True)
