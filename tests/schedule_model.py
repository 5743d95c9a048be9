"""Python restatement of the conflict-aware segment schedule of the packed-16 layout
(ccfindr_b200/csrc/kernels_common.cuh: class_steps, plan_class, make_schedule, schedule_item).
Test infrastructure: it documents the algorithm and lets its invariants be checked on the CPU; the
product builds the layout on the device."""
import numpy as np


def class_steps(n, L, NL, mult):
    k0 = max((n + NL - 1) // NL, (L + 1) // 2)
    return ((k0 + mult - 1) // mult) * mult


def plan_class(c, NL, K):
    """(R single steps, P pair steps), R + P = K, P minimal."""
    L = max(c)
    P = max(0, L - K)
    while True:
        R = K - P
        if sum(max(0, x - R) for x in c) <= NL * P:
            return R, P
        P += 1


def res_class(rr, NL):
    return 0 if NL == 8 else rr & 1


def res_bucket(rr, NL):
    return rr if NL == 8 else rr >> 1


def make_schedule(cnt8, NLp, kmult=4, sbs=False):
    """kmult (NL = 8 only): the step count is a multiple of it -- 4: every stored step is
    scheduled; 1: fewest steps, the rest of the last stored chunk of 4 steps is all holes.
    sbs (8 lanes, the 4-unit split layout): the parity classes are scheduled like the NL = 4
    classes but share their steps, even rows in lanes 0..3 and odd rows in lanes 4..7."""
    NL = 4 if sbs else NLp
    ncls, NB = (1, 8) if NL == 8 else (2, 4)
    sc = dict(K=[0, 0], pl=[(0, 0), (0, 0)], offr=[[0] * 8, [0] * 8], cnt=[[0] * 8, [0] * 8])
    for rr in range(8):
        sc["cnt"][res_class(rr, NL)][res_bucket(rr, NL)] = cnt8[rr]
    for cl in range(ncls):
        c = sc["cnt"][cl][:NB]
        sc["K"][cl] = class_steps(sum(c), max(c), NL, kmult if (NL == 8 or sbs) else 2)
    if sbs:
        sc["K"][0] = sc["K"][1] = max(sc["K"])
    elif ncls == 2 and ((sc["K"][0] + sc["K"][1]) & 3):
        sc["K"][1] += 2
    for cl in range(ncls):
        if sc["K"][cl] == 0:
            continue
        c = sc["cnt"][cl][:NB]
        sc["pl"][cl] = plan_class(c, NL, sc["K"][cl])
        acc = 0
        for b in range(NB):
            sc["offr"][cl][b] = acc
            acc += max(0, c[b] - sc["pl"][cl][0])
    return sc


def schedule_item(sc, NLp, rr, k, sbs=False):
    """Item index (step * NL + lane) of the k-th nonzero whose tile row has residue rr."""
    NL = 4 if sbs else NLp
    cl, b = res_class(rr, NL), res_bucket(rr, NL)
    R, P = sc["pl"][cl]
    if k < R:
        step, lane = k, b
    else:
        t = sc["offr"][cl][b] + (k - R)
        row, level = t % (2 * P), t // (2 * P)
        step = R + row % P
        lane = NL - 1 - level if row >= P else level
    if sbs:
        return step * 8 + cl * 4 + lane
    return ((sc["K"][0] if cl else 0) + step) * NL + lane


def wavefronts_sbs(cnt8, kmult=1):
    """4-unit split layout: (steps K, wavefronts of block A per unit, of block B per unit).
    Block A is conflict free when lanes 0..3 hold even rows and lanes 4..7 odd rows (asserted);
    block B costs the largest residue multiplicity of the step."""
    sc = make_schedule(cnt8, 8, kmult, sbs=True)
    K = sc["K"][0]
    mult = np.zeros((K, 8), dtype=int)
    seen = set()
    for rr in range(8):
        for k in range(cnt8[rr]):
            p = schedule_item(sc, 8, rr, k, sbs=True)
            assert 0 <= p < K * 8 and p not in seen
            seen.add(p)
            step, lane = divmod(p, 8)
            assert (lane >> 2) == (rr & 1)           # parity class <-> lane half
            mult[step, rr] += 1
    return K, K, int(np.maximum(mult.max(axis=1), 1).sum()) if K else 0


def wavefronts(cnt8, NL, kmult=4):
    """(steps, wavefronts per gather) of the schedule for residue counts cnt8: a step costs the
    largest number of its rows in one residue class (at least 1: hole steps still issue)."""
    sc = make_schedule(cnt8, NL, kmult)
    K = sc["K"][0] + sc["K"][1]
    mult = np.zeros((K, 8), dtype=int)
    seen = set()
    for rr in range(8):
        for k in range(cnt8[rr]):
            p = schedule_item(sc, NL, rr, k)
            assert 0 <= p < K * NL and p not in seen
            seen.add(p)
            mult[p // NL, rr] += 1
    if NL == 4:      # the four rows of a step must have equal parity
        for s in range(K):
            assert not (mult[s, 0::2].any() and mult[s, 1::2].any())
    return K, int(np.maximum(mult.max(axis=1), 1).sum())


def virtual_counts(cnt8):
    """kSchedCls4 (split layout with two dense units in block B, ranks 19 and 20): four classes
    (row mod 4) of two lanes each; the nonzeros of class c are dealt alternately to the virtual
    residues c and c + 4, which are scheduled like eight residue classes."""
    v = [0] * 8
    for b in range(4):
        c4 = cnt8[b] + cnt8[b + 4]
        v[b], v[b + 4] = (c4 + 1) // 2, c4 // 2
    return v


def wavefronts_cls4(rows, kmult=1):
    """(steps K, wavefronts of the two block-B gathers) for the tile rows of one segment.  Lane l of
    a step reads unit (c XOR (l >> 2)) of block B in gather c; a row of block B (32 bytes) covers
    the bank groups 2 (row mod 4) and 2 (row mod 4) + 1."""
    rows = list(rows)
    cnt8 = np.bincount(np.asarray(rows, dtype=int) % 8, minlength=8).tolist()
    sc = make_schedule(virtual_counts(cnt8), 8, kmult)
    K = sc["K"][0]
    bg = np.zeros((2, K, 8), dtype=int)
    seen = set()
    rank = {}
    # rank in the class: residues c first, then c + 4 (the device code's order)
    for rr in list(range(4)) + list(range(4, 8)):
        for row in [x for x in rows if x % 8 == rr]:
            c = row % 4
            kc = rank.get(c, 0)
            rank[c] = kc + 1
            p = schedule_item(sc, 8, c + 4 * (kc & 1), kc >> 1)
            assert 0 <= p < K * 8 and p not in seen
            seen.add(p)
            step, lane = divmod(p, 8)
            for g in range(2):
                bg[g, step, (2 * c + (g ^ (lane >> 2))) % 8] += 1
    w = int(np.maximum(bg.max(axis=2), 1).sum()) if K else 0
    return K, w


def make_schedule_one4(c4, kmult=1):
    """kSchedOne4 (4-lane groups): ONE set of four classes (row mod 4), one lane each; K =
    max(ceil(n / 4), ceil(L / 2)) steps, R single + P pair steps (make_schedule(..., one4))."""
    n, L = sum(c4), max(c4)
    K = class_steps(n, L, 4, kmult)
    sc = dict(K=[K, 0], pl=[(0, 0), (0, 0)], offr=[[0] * 8, [0] * 8], cnt=[list(c4) + [0] * 4, [0] * 8])
    if K:
        sc["pl"][0] = plan_class(list(c4), 4, K)
        acc = 0
        for b in range(4):
            sc["offr"][0][b] = acc
            acc += max(0, c4[b] - sc["pl"][0][0])
    return sc


def schedule_item_one4(sc, c, k):
    """Item index (step * 4 + lane) of the k-th nonzero of class c."""
    R, P = sc["pl"][0]
    if k < R:
        return k * 4 + c
    t = sc["offr"][0][c] + (k - R)
    row, level = t % (2 * P), t // (2 * P)
    return (R + row % P) * 4 + (3 - level if row >= P else level)


def wavefronts_one4_pair(rows0, rows1):
    """Two 4-lane groups share a bank phase (quarter-warp): group 0 (lanes 0..3) reads unit c of
    block B in gather c, group 1 (lanes 4..7) unit c XOR 1, so they use the even and the odd bank
    groups and never collide with each other.  Returns (steps K of the longer segment, wavefronts
    of the two block-B gathers); block A costs one wavefront per step and gather whatever the
    rows are."""
    plans = []
    for rows in (rows0, rows1):
        rows = list(rows)
        c4 = np.bincount(np.asarray(rows, dtype=int) % 4, minlength=4).tolist()
        sc = make_schedule_one4(c4)
        seen, place, rank = set(), {}, [0] * 4
        # rank in the class: residues c first, then c + 4 (the device code's order)
        for rr in range(8):
            for row in [x for x in rows if x % 8 == rr]:
                c = row % 4
                p = schedule_item_one4(sc, c, rank[c])
                rank[c] += 1
                assert 0 <= p < sc["K"][0] * 4 and p not in seen
                seen.add(p)
                place[p] = c
        plans.append((sc["K"][0], place))
    K = max(plans[0][0], plans[1][0])
    w = 0
    for g in range(2):                     # gather g reads unit g ^ h of block B
        for step in range(K):
            bg = np.zeros(8, dtype=int)
            for h, (Kh, place) in enumerate(plans):
                for lane in range(4):
                    c = place.get(step * 4 + lane)
                    if c is not None:
                        bg[(2 * c + (g ^ h)) % 8] += 1
            w += max(int(bg.max()), 1)
    return K, w


def window_order(steps, NO, W):
    """Stored order of the segments of one pass (seg_order_keys_kernel): inside every window of W
    consecutive owners of a slab, by decreasing number of steps; stable.  steps[e], e = slab * NO
    + owner.  Returns order[pos] = e."""
    steps = np.asarray(steps)
    E = len(steps)
    e = np.arange(E)
    slab, o = e // NO, e % NO
    nwin = (NO + W - 1) // W
    key = (slab * nwin + o // W) * 1024 + (1023 - np.minimum(steps, 1023))
    return np.argsort(key, kind="stable")
