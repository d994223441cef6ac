"""CPU: size-independent properties of the file-level schedule and the cut search (hypothesis)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from speechcatcher_b200.recognize import plan_segments


@settings(max_examples=150, deadline=None)
@given(n=st.integers(1, 40_000_000), chunk=st.sampled_from([4000, 8192, 16000, 25600]),
       cuts=st.lists(st.integers(2000, 18000), min_size=0, max_size=40))
def test_schedule_covers_every_sample_once(n, chunk, cuts):
    """Whatever the segmentation: the calls of all segments tile [0, n) in order without gaps or overlap, every segment
    ends with exactly one final call, and only the very last call of the file finalises everything."""
    ends = np.cumsum(cuts).tolist()
    segments = list(zip([0] + ends[:-1], ends))
    plan = plan_segments(n, 16000, segments, chunk)
    assert plan.finalize_iters[0] == -1 and plan.finalize_iters[-1] == plan.max_i == n // chunk + 1
    assert len(plan.seconds) == plan.n_segments and plan.seconds[0][0] == 0 and plan.seconds[-1][1] == n / 16000
    its = plan.finalize_iters
    if any(b < a for a, b in zip(its, its[1:])):
        return                      # boundaries closer than one chunk collapse; the reference has the same degenerate case
    pos, n_all = 0, 0
    for k in range(plan.n_segments):
        calls = plan.calls(k, n, chunk)
        for j, (a, b, fin, fin_all) in enumerate(calls):
            assert a == min(pos, n) and a <= b <= n and b - a <= chunk
            pos = b if b > a else pos
            assert fin == (j == len(calls) - 1)
            n_all += int(fin_all)
            if fin_all:
                assert fin and k == plan.n_segments - 1 and a == b        # the last call of a file is an empty final chunk
        # segment times are consistent with its first call (within one chunk)
        if calls:
            assert abs(plan.seconds[k][0] * 16000 - calls[0][0]) <= chunk + 160
    assert pos == n and n_all == 1


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(50, 30_000), beam=st.integers(1, 6), step=st.integers(1, 25),
       min_len=st.integers(0, 400), look=st.integers(100, 3000), ideal=st.integers(50, 2000))
def test_cut_search_properties(seed, n, beam, step, min_len, look, ideal):
    from speechcatcher_b200.simple_endpointing import BeamSearch
    rng = np.random.default_rng(seed)
    curve = rng.standard_normal(n) * 0.5 - 0.2
    segs = BeamSearch(beam_size=beam, ideal_segment_len=ideal, max_lookahead=look, min_len=min_len, step=step,
                      len_reward_weight=1.5, energy_weight=1.0).search(curve, n)
    assert segs and segs[0][0] == 0
    for (a, b), (c, d) in zip(segs, segs[1:]):
        assert b == c
    if segs[-1][1] != n or len(segs) > 1:
        for a, b in segs:
            assert min_len + 1 <= b - a <= look and (b - a - 1 - min_len) % step == 0 and b <= n - 1
