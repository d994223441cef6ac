"""GPU: file-level recognition (SURVEY.md 8(f) N1 + N2) -- segments of one file as concurrent streams of one
StreamGroup, against the same schedule driven through CPU oracles."""
import numpy as np
import pytest

from helpers import OracleGroup, model_dir

pytestmark = pytest.mark.gpu


def _strip(aux):
    return [{k: v for k, v in p.items()} for p in aux]


@pytest.mark.parametrize("sharpen,eos_bias", [(1.0, 0.0), (1.0, 6.0), (1.0, 8.0)])
def test_segments_as_streams_match_oracle(sharpen, eos_bias):
    from speechcatcher_b200 import Speech2TextStreaming, StreamGroup
    from speechcatcher_b200.recognize import recognize
    from speechcatcher_b200.synthetic import synth_audio
    md = model_dir("m_d2", sharpen=sharpen, eos_bias=eos_bias)
    audio = (synth_audio(3, 30 * 16000 + 321) * 32768.0).astype(np.int16)
    segs = [(0, 800), (800, 1700)]
    want_text, want_aux = recognize(OracleGroup(md, 3), audio, 16000, segments=segs)
    assert len(want_aux) >= 1 and sum(len(p["tokens"]) for p in want_aux) > 20
    if eos_bias:        # hypotheses end regularly: every segment inside the file contributes text
        stamps = want_aux[0]["token_timestamps"]
        assert any(8.0 <= t < 17.0 for t in stamps) and max(stamps) > 17.0 and (eos_bias < 8.0 or min(stamps) < 8.0)
    # three segments in lock step on one engine
    group = StreamGroup(md, n_streams=3, beam_size=5, max_seconds=40.0)
    text, aux = recognize(group, audio, 16000, segments=segs)
    assert text == want_text and _strip(aux) == _strip(want_aux)
    # fewer streams than segments: a freed stream picks up the waiting segment
    group2 = StreamGroup(md, n_streams=2, beam_size=5, max_seconds=40.0)
    text2, aux2 = recognize(group2, audio, 16000, segments=segs)
    assert text2 == want_text and aux2 == want_aux
    # the drop-in facade: serial like the reference (num_processes=1), then grown to 3 streams (num_processes=3)
    s2t = Speech2TextStreaming(md, beam_size=5, device="cuda:0")
    text3, aux3 = recognize(s2t, audio, 16000, num_processes=1, segments=segs)
    assert text3 == want_text and aux3 == want_aux
    text4, aux4 = recognize(s2t, audio, 16000, num_processes=3, segments=segs)
    assert s2t.group.n_streams == 3
    assert text4 == want_text and aux4 == want_aux


def test_long_file_with_device_segmentation_matches_oracle():
    """> 60 s: the device segmenter picks the cuts, the segments decode as streams; the checker is the oracle chain."""
    from oracle.endpointing import segment_speech_oracle
    from oracle.gen_golden_endpointing import pause_audio
    from speechcatcher_b200 import StreamGroup
    from speechcatcher_b200.recognize import recognize
    md = model_dir("m_d2", eos_bias=7.0)
    audio = pause_audio(21, 64.0)
    kw = dict(average_segment_length=20.0, max_segment_len_sec=30)

    def dev_segmenter(data, rate):
        from speechcatcher_b200.simple_endpointing import segment_speech
        return segment_speech(data, rate, **kw)

    cpu_segs = segment_speech_oracle(audio, 16000, **kw)
    assert len(cpu_segs) >= 2
    want = recognize(OracleGroup(md, 4), audio, 16000, segmenter=lambda d, r: cpu_segs)
    got = recognize(StreamGroup(md, n_streams=4, beam_size=5, max_seconds=40.0), audio, 16000, segmenter=dev_segmenter)
    assert got == want
