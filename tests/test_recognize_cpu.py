"""CPU: the file-level glue against goldens produced by the reference's own `recognize`
(oracle/gen_golden_recognize.py), using the scripted recogniser on both sides."""
import json

import pytest

from helpers import GOLDEN
from oracle.gen_golden_recognize import case_audio
from oracle.scripted_backend import ScriptedGroup
from speechcatcher_b200.recognize import linear_interpolate_pos, plan_segments, recognize

G = json.loads((GOLDEN / "recognize.json").read_text())


@pytest.mark.parametrize("n_streams", [1, 3, 16])
@pytest.mark.parametrize("case", G["recognize"], ids=lambda c: c["name"])
def test_recognize_matches_reference_golden(case, n_streams):
    a = case_audio(case["seed"], case["n"])
    group = ScriptedGroup(n_streams)
    text, aux = recognize(group, a, 16000, chunk_length=case["chunk"], num_processes=n_streams, progress=False,
                          quiet=True, segments=[tuple(s) for s in (case["segments"] or [])])
    assert text == case["text"]
    assert aux == case["aux"]
    # every stream was reset after each of its segments (plus the initial reset)
    plan = plan_segments(case["n"], 16000, case["segments"] or [], case["chunk"])
    assert sum(s.n_resets for s in group.streams) == 2 * n_streams + plan.n_segments
    if n_streams == 1:      # serial order = the reference's call sequence: (samples, is_final, finalize_all)
        calls = [(n, f) for push in group.pushes for _, n, f in push]
        assert calls == [(n, f) for n, f, _ in case["calls"]]
    else:                   # lock step: concurrent segments share pushes, never more streams than the group holds
        assert max(len(p) for p in group.pushes) <= n_streams
        assert sum(len(p) for p in group.pushes) == len(case["calls"])


def test_segmenter_hook_is_used_only_for_long_files():
    seen = []

    def segmenter(data, rate):
        seen.append(len(data))
        return [(0, 3000), (3000, 7000)]
    recognize(ScriptedGroup(2), case_audio(1, 59 * 16000), 16000, segmenter=segmenter)
    assert seen == []
    text, aux = recognize(ScriptedGroup(2), case_audio(1, 100 * 16000), 16000, segmenter=segmenter)
    assert seen == [100 * 16000] and aux[0]["start"] == 0 and aux[-1]["end"] == 100.0


def test_capacity_is_checked_before_decoding():
    with pytest.raises(ValueError, match="max_seconds"):
        recognize(ScriptedGroup(2, max_seconds=61.0), case_audio(1, 100 * 16000), 16000, segments=[])
    with pytest.raises(ValueError, match="native"):
        recognize(ScriptedGroup(1), case_audio(1, 16000), 16000, decoder_impl="espnet")


@pytest.mark.parametrize("case", G["interp"], ids=lambda c: str(c["inp"]))
def test_linear_interpolate_pos_matches_reference(case):
    assert linear_interpolate_pos(case["inp"]) == case["out"]
