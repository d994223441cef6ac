"""CPU: live-session glue against goldens produced by the reference's own
SpeechRecognitionSession (oracle/gen_golden_live.py), with the scripted recogniser on both sides."""
import json
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import GOLDEN
from oracle.gen_golden_live import chunk_script
from oracle.scripted_backend import ScriptedLive
from speechcatcher_b200.live import LiveEndpointer, LiveSession, process_many

G = json.loads((GOLDEN / "live.json").read_text())


@pytest.mark.parametrize("case", G, ids=lambda c: c["name"])
def test_live_session_matches_reference_golden(case):
    backend = ScriptedLive()
    sess = LiveSession(backend, finalize_update_iters=case["finalize_update_iters"],
                       max_partial_iters=case["max_partial_iters"], vosk_output_format=case["vosk"])
    outputs = [sess.process_audio_chunk(c) for c in chunk_script(case["seed"], case["n"], "messages" in case["name"])]
    assert json.loads(json.dumps(outputs)) == case["outputs"]
    assert [list(c) for c in backend.log] == case["calls"]          # same samples (fp16-rounded) and final flags


def test_s16le_bytes_equal_int16_arrays():
    a, b = ScriptedLive(), ScriptedLive()
    sa, sb = LiveSession(a, vosk_output_format=True), LiveSession(b, vosk_output_format=True)
    for c in chunk_script(9, 30, False):
        assert sa.process_audio_chunk(c) == sb.process_audio_chunk(c.tobytes())
    with pytest.raises(NotImplementedError):
        LiveSession(a, audio_format="webm")


def test_microphone_variant_of_the_rule():
    """speechcatcher.py:714-722: at least 7 lengths, then finalise when the last TEN are equal; no iteration cap."""
    rule = LiveEndpointer(7, max_iters=None, window=10)
    lens = [1, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3]
    fired = []
    for x in lens:
        fired.append(rule.decide())
        rule.observe(x)
    ref, n_best, want = 7, [], []
    for x in lens:                                    # the loop body of the reference, restated inline
        if len(n_best) < ref:
            f = False
        elif all(v == n_best[-1] for v in n_best[-10:]):
            f, n_best = True, []
        else:
            f = False
        want.append(f)
        n_best += [x]
    assert fired == want and any(want)


class _FakeGroup:
    """StreamGroup protocol over scripted live recognisers (one per stream)."""

    def __init__(self, n):
        self.rec = [ScriptedLive() for _ in range(n)]
        self.last, self.n_pushes = {}, 0

    def push(self, ids, chunks, fins):
        self.n_pushes += 1
        for s, c, f in zip(ids, chunks, fins):
            self.last[s] = self.rec[s](c, is_final=f)

    def last_plan(self, s):
        return SimpleNamespace(called=1)

    def beam(self, s):
        return None

    def results(self, s, is_final, finalize_all, token_list=None):
        return self.last[s]


def test_process_many_equals_per_session_calls():
    n = 5
    group = _FakeGroup(n)
    views = [SimpleNamespace(group=group, stream_id=k, token_list=None, _calls_since_reset=0, beam_state=None,
                             reset=lambda: None) for k in range(n)]
    batched = [LiveSession(v, finalize_update_iters=3, vosk_output_format=True) for v in views]
    solo_rec = [ScriptedLive() for _ in range(n)]
    solo = [LiveSession(r, finalize_update_iters=3, vosk_output_format=True) for r in solo_rec]
    scripts = [chunk_script(20 + k, 40, True) for k in range(n)]
    for t in range(40):
        got = process_many(batched, [scripts[k][t] for k in range(n)])
        want = [solo[k].process_audio_chunk(scripts[k][t]) for k in range(n)]
        assert got == want, t
    assert group.n_pushes <= 40                        # one batched push per step, not one per session
    assert [r.log for r in group.rec] == [r.log for r in solo_rec]


def test_best_hypothesis_partials_drive_the_rule():
    """partials="best": the rule sees the running hypothesis grow and stall (silence), instead of empty partials."""
    class Rec:
        def __init__(self):
            self.n, self.log = 0, []

        def reset(self):
            self.n = 0

        def __call__(self, speech, is_final=False, finalize_all=False):
            self.log.append(bool(is_final))
            if float(np.abs(np.asarray(speech, np.float32)).max()) > 0.01:
                self.n += 1
            out = [("x" * self.n, ["x"] * self.n, [5] * self.n, list(range(self.n)), {})] if is_final else []
            if is_final:
                self.n = 0
            return out

        def get_best_hypothesis(self):
            return ("x" * self.n, ["x"] * self.n, [5] * self.n, list(range(self.n)), {}) if self.n else None

    loud = (np.ones(4096) * 3000).astype(np.int16)
    quiet = np.zeros(4096, np.int16) + 1
    rec = Rec()
    sess = LiveSession(rec, finalize_update_iters=3, partials="best")
    outs = [sess.process_audio_chunk(c) for c in [loud] * 5 + [quiet] * 4 + [loud] * 2]
    assert outs[:5] == ["x", "xx", "xxx", "xxxx", "xxxxx"]
    assert rec.log.index(True) == 7                    # lengths 1..5, then 5, 5: the last three are equal -> finalise
    assert outs[7] == "xxxxx.\n" and outs[8] == "" and outs[9] == "x"
    ref_mode = Rec()
    sess2 = LiveSession(ref_mode, finalize_update_iters=3)           # reference partials: nothing to observe
    assert [sess2.process_audio_chunk(c) for c in [loud] * 5] == [""] * 5 and True not in ref_mode.log
    with pytest.raises(ValueError):
        LiveSession(rec, partials="all")
