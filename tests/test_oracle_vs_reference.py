"""Live comparison of the oracle with the reference package, when /root/reference is present (build container).
On the GPU box these tests skip; the committed golden vectors carry the same information."""
import numpy as np
import pytest

import oracle


@pytest.mark.parametrize("length,tw", ((400, 2), (200, 1), (100, 0.5)))
def test_de_psd_live(reference, length, tw):
    rng = np.random.default_rng(length)
    x = (30 * rng.standard_normal((62, length)) + rng.uniform(-50, 50, (62, 1))).astype(np.float32)
    de_ref, psd_ref = reference["DE_PSD"].DE_PSD(x, 200, tw)
    de, psd = oracle.de_psd_loop(x, 200, tw)
    assert np.array_equal(de, de_ref) and np.array_equal(psd, psd_ref)
    de, psd = oracle.de_psd_closed_form(x, 200, tw)
    assert np.max(np.abs(de - de_ref)) < 1e-12 and np.max(np.abs(psd - psd_ref) / psd_ref) < 1e-12


def test_segmentation_live(reference):
    rng = np.random.default_rng(1)
    raw = rng.integers(-1000, 1000, (7, 3, 104000)).astype(np.int16)
    seg = reference["segment_raw_signals_200Hz"].extract_2s_segment
    for b, c, r in ((0, 0, 0), (6, 39, 4), (2, 11, 3)):
        assert np.array_equal(seg(block=b, concept=c, repetition=r, data=raw),
                              oracle.extract_2s_segment(block=b, concept=c, repetition=r, data=raw))


def test_sliding_window_live(reference):
    rng = np.random.default_rng(2)
    clips = rng.standard_normal((2, 3, 5, 6, 400)).astype(np.float32)
    ref = reference["segment_sliding_window"].seg_sliding_window(clips, 0.5, 0.25, fs=200)
    assert np.array_equal(ref, oracle.seg_sliding_window(clips, 0.5, 0.25, fs=200))


def test_drivers_live(reference):
    rng = np.random.default_rng(3)
    clips = (30 * rng.standard_normal((1, 2, 2, 62, 400))).astype(np.float32)
    de_ref, psd_ref = reference["extract_DE_PSD_features_1per2s"].extract_de_psd_raw(clips, 200)
    de, psd = oracle.extract_de_psd_raw(clips, 200, closed=False)
    assert np.array_equal(de, de_ref) and np.array_equal(psd, psd_ref)
    win = np.ascontiguousarray(oracle.seg_sliding_window(clips, 0.5, 0.25))
    de_ref, psd_ref = reference["extract_DE_PSD_features_1per500ms"].extract_de_psd_sw(win, 200, 0.5)
    de, psd = oracle.extract_de_psd_sw(win, 200, 0.5, closed=False)
    assert np.array_equal(de, de_ref) and np.array_equal(psd, psd_ref)


def test_seq2seq_windows_live(reference):
    """The Seq2Seq trainer's inline loop, executed from the reference file by line number."""
    import importlib.util
    import os
    import torch
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_seq2seq",
                                                  os.path.join(here, "golden", "make_golden_seq2seq.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 2, 3, 4, 400)).astype(np.float32)
    ref = mod.reference_windows(torch.from_numpy(x)).numpy()
    assert np.array_equal(ref, oracle.seq2seq_windows(x))
