"""The oracle against the golden vectors produced by the reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle
from conftest import decode

LENGTHS = ((400, 2), (200, 1), (100, 0.5))
KINDS = ("white", "offset", "pink_tone", "small")


@pytest.mark.parametrize("length,tw", LENGTHS)
@pytest.mark.parametrize("kind", KINDS)
def test_loop_oracle_is_bit_exact_with_reference(golden, length, tw, kind):
    g = golden("de_psd_golden.npz")
    key = f"L{length}_{kind}"
    de, psd = oracle.de_psd_loop(decode(g[key + "_codes"]), 200, tw)
    assert np.array_equal(de, g[key + "_de"])
    assert np.array_equal(psd, g[key + "_psd"])


@pytest.mark.parametrize("length,tw", LENGTHS)
@pytest.mark.parametrize("kind", KINDS)
def test_closed_form_matches_reference(golden, length, tw, kind):
    g = golden("de_psd_golden.npz")
    key = f"L{length}_{kind}"
    de, psd = oracle.de_psd_closed_form(decode(g[key + "_codes"]), 200, tw)
    assert np.max(np.abs(psd - g[key + "_psd"]) / g[key + "_psd"]) < 1e-12
    assert np.max(np.abs(de - g[key + "_de"])) < 1e-12


def test_return_order_is_de_then_psd(golden):
    g = golden("de_psd_golden.npz")
    de, psd = oracle.de_psd_loop(decode(g["L200_white_codes"]), 200, 1)
    assert np.allclose(de, np.log2(100 * psd))


def test_impulse_known_answer(golden):
    """A unit impulse at sample 0 has |X[k]|^2 = (a h[0])^2 in every bin, so every band mean equals it."""
    g = golden("de_psd_golden.npz")
    h0 = oracle.hann_window(100)[0]
    amp = g["impulse_x"][:, 0].astype(np.float64)
    expect = (amp * h0) ** 2
    assert np.allclose(g["impulse_psd"], expect[:, None], rtol=1e-12)
    de, psd = oracle.de_psd_closed_form(g["impulse_x"], 200, 0.5)
    assert np.allclose(psd, expect[:, None], rtol=1e-12)
    assert np.allclose(de, g["impulse_de"], atol=1e-12)


def test_band_ranges_and_hann():
    assert oracle.band_bin_ranges(200) == [(0, 3), (3, 7), (7, 13), (13, 30), (30, 98)]
    h = oracle.hann_window(100)
    assert h.shape == (100,) and h.dtype == np.float64
    assert np.allclose(h, h[::-1]) and h[0] == 0.5 - 0.5 * np.cos(2 * np.pi / 101)


def test_two_second_mode_ignores_second_half():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 400)).astype(np.float32)
    y = x.copy()
    y[:, 200:] = 1e6
    assert np.array_equal(oracle.de_psd_loop(x, 200, 2)[1], oracle.de_psd_loop(y, 200, 2)[1])


def test_length_mismatch_and_zero_power_errors():
    with pytest.raises(ValueError):
        oracle.de_psd_loop(np.ones((2, 150), np.float32), 200, 1)
    with pytest.raises(ValueError, match="math domain error"):
        oracle.de_psd_loop(np.zeros((2, 200), np.float32), 200, 1)


def _formula_raw(n_ch, n_t):
    b = np.arange(7).reshape(7, 1, 1)
    ch = np.arange(n_ch).reshape(1, n_ch, 1)
    t = np.arange(n_t).reshape(1, 1, n_t)
    return (((b * 7919 + ch * 104729 + t * 31) % 65536) - 32768).astype(np.int16)


def test_segmentation_golden(golden):
    g = golden("segment_golden.npz")
    raw = _formula_raw(int(g["n_ch"]), int(g["n_t"]))
    for (b, c, r), seg in zip(g["picks"], g["segments"]):
        got = oracle.extract_2s_segment(block=int(b), concept=int(c), repetition=int(r), data=raw)
        assert np.array_equal(got, seg)
        assert oracle.clip_start(int(c), int(r)) == int(c) * 2600 + 600 + int(r) * 400
    full = oracle.segment_subject(raw)
    assert full.shape == (7, 40, 5, int(g["n_ch"]), 400) and full.dtype == np.int16
    for (b, c, r), seg in zip(g["picks"], g["segments"]):
        assert np.array_equal(full[b, c, r], seg)


def test_segmentation_errors():
    raw = np.zeros((7, 2, 104000), np.float32)
    for kw in (dict(block=7, concept=0, repetition=0), dict(block=0, concept=40, repetition=0),
               dict(block=0, concept=0, repetition=5), dict(block=-1, concept=0, repetition=0)):
        with pytest.raises(ValueError):
            oracle.extract_2s_segment(data=raw, **kw)
    with pytest.raises(RuntimeError, match="Segment length mismatch"):
        oracle.extract_2s_segment(block=0, concept=39, repetition=4, data=raw[:, :, :103999])


def test_drivers_golden(golden):
    g = golden("drivers_golden.npz")
    clips = decode(g["codes"])
    de, psd = oracle.extract_de_psd_raw(clips, 200, closed=False)
    assert de.dtype == np.float32 and np.array_equal(de, g["de_2s"]) and np.array_equal(psd, g["psd_2s"])
    de, psd = oracle.extract_de_psd_1s(clips, 200, closed=False)
    assert de.dtype == np.float64 and np.array_equal(de, g["de_1s"]) and np.array_equal(psd, g["psd_1s"])
    win = oracle.seg_sliding_window(clips, 0.5, 0.25, fs=200)
    assert win.shape == tuple(g["window_shape"])
    de, psd = oracle.extract_de_psd_sw(win, 200, 0.5, closed=False)
    assert de.dtype == np.float32 and np.array_equal(de, g["de_500ms"]) and np.array_equal(psd, g["psd_500ms"])
    # vectorised form of the drivers: same numbers to float32 rounding
    de_c, psd_c = oracle.extract_de_psd_sw(win, 200, 0.5, closed=True)
    assert np.allclose(psd_c, g["psd_500ms"], rtol=1e-6) and np.allclose(de_c, g["de_500ms"], atol=1e-5)


def test_one_second_script_golden(golden):
    """Sampled outputs of the reference's real 1 s script (run with runpy on a full-size subject)."""
    g = golden("script_1s_golden.npz")
    assert tuple(g["out_shape"]) == (7, 40, 5, 2, 62, 5) and str(g["out_dtype"]) == "float64"
    clips = decode(g["clips_codes"])                      # (n, 62, 400)
    de, psd = oracle.extract_de_psd_1s(clips[None, None], 200, closed=False)
    assert np.array_equal(de[0, 0], g["de"]) and np.array_equal(psd[0, 0], g["psd"])


def test_seq2seq_window_layout_golden(golden):
    """oracle.seq2seq_windows against the output of the reference's own lines
    (my_autoregressive_transformer.py:309-314, executed by tests/golden/make_golden_seq2seq.py)."""
    g = golden("seq2seq_golden.npz")
    got = oracle.seq2seq_windows(g["codes"].astype(np.float32))
    assert got.shape == (2, 3, 5, 100, 7)
    assert np.array_equal(got, g["windows"].astype(np.float32))
    # the same windows as seg_sliding_window (window axis before the channel axis there, last here)
    sw = oracle.seg_sliding_window(g["codes"][None].astype(np.float32), 0.5, 0.25)[0]       # (2, 3, 7, 5, 100)
    assert np.array_equal(np.moveaxis(sw, 2, -1), got)
