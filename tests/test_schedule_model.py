"""Invariants of the conflict-aware segment schedule (tests/schedule_model.py restates the device
code of ccfindr_b200/csrc/kernels_common.cuh): every nonzero gets its own slot, the steps are whole
chunks, and the wavefront count is the optimum max(K, largest bucket) for the K steps executed."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from schedule_model import (make_schedule, wavefronts, wavefronts_sbs, wavefronts_cls4,
                            wavefronts_one4_pair, window_order)

counts = st.lists(st.integers(min_value=0, max_value=90), min_size=8, max_size=8).filter(lambda c: sum(c) > 0)


@settings(max_examples=300, deadline=None)
@given(counts)
def test_one_lane_per_nonzero_schedule_is_collision_free_and_optimal(cnt):
    K, w = wavefronts(cnt, 8)
    R, P = make_schedule(cnt, 8)["pl"][0]
    assert K % 4 == 0 and K * 8 >= sum(cnt) and R + P == K
    # never below the bound, never above single steps + 2-way pair steps; P is the smallest number
    # of pair steps whose 8 P slots hold what does not fit into R single steps
    assert max(K, max(cnt)) <= w <= R + 2 * P
    assert P == 0 or max(cnt) > R + 2 * P - 1 or sum(max(0, c - (R + 1)) for c in cnt) > 8 * (P - 1)


@settings(max_examples=300, deadline=None)
@given(counts)
def test_two_lanes_per_nonzero_schedule(cnt):
    K, w = wavefronts(cnt, 4)
    sc = make_schedule(cnt, 4)
    assert K % 4 == 0 and K == sc["K"][0] + sc["K"][1]
    even, odd = cnt[0::2], cnt[1::2]
    lo = max(sc["K"][0], max(even)) + max(sc["K"][1], max(odd))
    hi = sum(R + 2 * P for R, P in sc["pl"])
    assert lo <= w <= max(hi, lo)


@pytest.mark.parametrize("cnt", [[9, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 1],
                                 [83, 79, 68, 53, 74, 61, 74, 195], [48, 50, 48, 48, 52, 39, 55, 36]])
def test_skewed_segments(cnt):
    for NL in (8, 4):
        K, w = wavefronts(cnt, NL)
        assert K % 4 == 0


def test_typical_segment_beats_round_robin():
    """~205 nonzeros per segment (C2): the schedule needs ~1.25 wavefronts per ideal wavefront,
    the round-robin order of the 8-byte layouts ~1.4 (bank conflicts)."""
    rng = np.random.default_rng(0)
    tot_w = tot_ideal = 0
    for _ in range(400):
        rows = rng.choice(2880, size=rng.binomial(2880, 0.072), replace=False)
        cnt = np.bincount(rows % 8, minlength=8).tolist()
        _, w = wavefronts(cnt, 8)
        tot_w += w
        tot_ideal += sum(cnt) / 8
    assert 1.15 < tot_w / tot_ideal < 1.32


@settings(max_examples=300, deadline=None)
@given(counts)
def test_fewest_steps_schedule_for_the_split_layout(cnt):
    """kmult = 1 (kernels that skip the gathers of hole entries): the schedule uses
    max(ceil(n/8), ceil(L/2)) steps, still collision free, and every nonzero sits inside them (the
    rest of the stored chunk of 4 steps is all holes)."""
    K, w = wavefronts(cnt, 8, kmult=1)
    sc = make_schedule(cnt, 8, kmult=1)
    R, P = sc["pl"][0]
    n, L = sum(cnt), max(cnt)
    assert K == max((n + 7) // 8, (L + 1) // 2) and R + P == K
    assert max(K, L) <= w <= R + 2 * P


def test_split_layout_cost_model():
    """Wavefronts per segment at r = 20 (10 units per row, ~105 nonzeros per segment): block A
    (8 units) costs one wavefront per executed step whatever the rows, block B (2 units) follows
    the residue schedule.  Against the lock-step layout (all 10 units pay the schedule)."""
    rng = np.random.default_rng(1)
    old = new = ideal = 0.0
    for _ in range(400):
        rows = rng.choice(1312, size=rng.binomial(1312, 0.08), replace=False)
        cnt = np.bincount(rows % 8, minlength=8).tolist()
        _, w4 = wavefronts(cnt, 8, kmult=4)
        K1, w1 = wavefronts(cnt, 8, kmult=1)
        old += 10 * w4
        new += 8 * K1 + 2 * w1
        ideal += 10 * sum(cnt) / 8
    assert old / ideal > 1.30 and new / ideal < 1.15


@settings(max_examples=300, deadline=None)
@given(counts)
def test_side_by_side_parity_schedule_of_the_four_unit_split(cnt):
    """Every nonzero gets its own slot in the lane half of its row parity; the step count is the
    larger class's max(ceil(n/4), ceil(L/2)); block B never sees more than a 2-way conflict."""
    K, wa, wb = wavefronts_sbs(cnt)
    even, odd = cnt[0::2], cnt[1::2]
    need = max(max((sum(c) + 3) // 4, (max(c) + 1) // 2) for c in (even, odd))
    assert K == need and wa == K and K <= wb <= 2 * K


def test_four_unit_split_cost_model():
    """Wavefronts per segment at r = 10 (5 units per row, ~230 nonzeros per segment): block A
    (4 units) costs one wavefront per step under the parity rule, block B (1 unit) follows the
    residues.  Against the lock-step layout where all 5 units pay max(K, largest bucket)."""
    rng = np.random.default_rng(2)
    old = new = ideal = 0.0
    for _ in range(400):
        rows = rng.choice(2880, size=rng.binomial(2880, 0.08), replace=False)
        cnt = np.bincount(rows % 8, minlength=8).tolist()
        _, w4 = wavefronts(cnt, 8, kmult=4)
        K, wa, wb = wavefronts_sbs(cnt)
        old += 5 * w4
        new += 4 * wa + 1 * wb
        ideal += 5 * sum(cnt) / 8
    assert old / ideal > 1.22 and new / ideal < 1.13


def test_four_classes_of_two_lanes_cost_model():
    """r = 20 with a dense block B (rows of exactly 160 bytes, T = 1440 instead of 1312): block A
    costs one wavefront per step, block B is scheduled in four classes of two lanes.  Against the
    eight residue classes of the padded block B (stride 3 units)."""
    rng = np.random.default_rng(3)
    old = new = ideal_old = ideal_new = 0.0
    for _ in range(400):
        rows = rng.choice(1312, size=rng.binomial(1312, 0.08), replace=False)
        cnt = np.bincount(rows % 8, minlength=8).tolist()
        K1, w1 = wavefronts(cnt, 8, kmult=1)
        old += 8 * K1 + 2 * w1
        ideal_old += 10 * len(rows) / 8
        rows = rng.choice(1440, size=rng.binomial(1440, 0.08), replace=False)
        K, wb = wavefronts_cls4(rows)
        assert K == max((len(rows) + 7) // 8, 1) or K > (len(rows) + 7) // 8
        new += 8 * K + wb
        ideal_new += 10 * len(rows) / 8
    assert old / ideal_old > 1.09 and new / ideal_new < 1.08


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=1439), min_size=1, max_size=300, unique=True))
def test_four_class_schedule_places_every_nonzero_once(rows):
    K, wb = wavefronts_cls4(rows)
    assert K * 8 >= len(rows) and 2 * K <= wb <= 8 * K


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=1439), min_size=1, max_size=200, unique=True),
       st.lists(st.integers(min_value=0, max_value=1439), min_size=1, max_size=200, unique=True))
def test_four_lane_groups_share_a_bank_phase_without_colliding(rows0, rows1):
    """Every nonzero of both segments has its own slot, and a step costs at most a 2-way conflict
    (the pair steps of ONE group): the two groups use disjoint bank groups."""
    K, w = wavefronts_one4_pair(rows0, rows1)
    assert K >= max((len(rows0) + 3) // 4, (len(rows1) + 3) // 4)
    assert 2 * K <= w <= 4 * K


def test_four_lane_groups_cost_model():
    """r = 20, T = 1440, ~115 nonzeros per segment: 4-lane groups store ceil(n / 4) steps (fewer
    hole slots than ceil(n / 8) * 8) and block B stays nearly conflict free."""
    rng = np.random.default_rng(4)
    tot = ideal = slots = nnz = 0.0
    for _ in range(300):
        r0 = rng.choice(1440, size=rng.binomial(1440, 0.08), replace=False)
        r1 = rng.choice(1440, size=rng.binomial(1440, 0.08), replace=False)
        K, wb = wavefronts_one4_pair(r0, r1)
        tot += 8 * K + wb                  # 8 conflict-free gathers of block A + 2 of block B
        ideal += 10 * max(len(r0), len(r1)) / 4
        slots += 8 * K
        nnz += len(r0) + len(r1)
    assert tot / ideal < 1.09
    assert slots / nnz < 1.12              # (unequal partners included)


def test_window_order_keeps_locality_and_groups_equal_lengths():
    rng = np.random.default_rng(5)
    NO, S, W = 1000, 3, 128
    steps = rng.integers(20, 40, size=NO * S)
    order = window_order(steps, NO, W)
    assert sorted(order.tolist()) == list(range(NO * S))           # a permutation
    pos = np.arange(NO * S)
    slab_of_pos, slab_of_e = pos // NO, order // NO
    assert (slab_of_pos == slab_of_e).all()                         # slabs keep their ranges
    o = order % NO
    assert (np.abs(o - pos % NO) < W).all()                         # owners stay inside their window
    # inside a window the steps are non-increasing
    for s0 in range(S):
        for w0 in range(0, NO, W):
            seg = steps[order[s0 * NO + w0: s0 * NO + min(w0 + W, NO)]]
            assert (np.diff(seg) <= 0).all()
