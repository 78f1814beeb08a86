"""Host-side logic of the reference-named shims: views, index arithmetic, exceptions.  CPU only."""
import os

import numpy as np
import pytest
import torch

import oracle
from eeg2video_b200 import cohort
from eeg2video_b200.EEG_preprocessing import segment_raw_signals_200Hz as seg
from eeg2video_b200.EEG_preprocessing import segment_sliding_window as ssw
from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per500ms as d500
from eeg2video_b200.EEG_preprocessing.DE_PSD import DE_PSD
from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per2s import extract_de_psd_raw
from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per1s import extract_de_psd_1s


def test_extract_2s_segment_is_a_view_and_bit_exact():
    rng = np.random.default_rng(0)
    raw = rng.integers(-999, 999, (7, 4, 104000)).astype(np.int16)
    for b, c, r in ((0, 0, 0), (6, 39, 4), (3, 20, 2)):
        got = seg.extract_2s_segment(block=b, concept=c, repetition=r, data=raw)
        assert got.shape == (4, 400) and got.dtype == raw.dtype
        assert np.shares_memory(got, raw)
        assert np.array_equal(got, oracle.extract_2s_segment(block=b, concept=c, repetition=r, data=raw))
    t = torch.from_numpy(raw)
    got = seg.extract_2s_segment(block=1, concept=2, repetition=3, data=t)
    assert got.data_ptr() == t[1, 0, 2 * 2600 + 600 + 3 * 400:].data_ptr()
    assert seg.__all__ == ["extract_2s_segment", "segment_all_files"] and seg.FS == 200


def test_extract_2s_segment_errors(tmp_path):
    raw = np.zeros((7, 2, 104000), np.float32)
    with pytest.raises(TypeError):
        seg.extract_2s_segment(0, 0, 0)                      # keyword-only
    with pytest.raises(ValueError, match=r"`block` must be in \[0, 6\]"):
        seg.extract_2s_segment(block=7, concept=0, repetition=0, data=raw)
    with pytest.raises(ValueError, match=r"`concept` must be in \[0, 39\]"):
        seg.extract_2s_segment(block=0, concept=40, repetition=0, data=raw)
    with pytest.raises(ValueError, match=r"`repetition` must be in \[0, 4\]"):
        seg.extract_2s_segment(block=0, concept=0, repetition=5, data=raw)
    with pytest.raises(ValueError, match="`subject` must be >= 1"):
        seg.extract_2s_segment(block=0, concept=0, repetition=0)
    with pytest.raises(FileNotFoundError):
        seg.extract_2s_segment(block=0, concept=0, repetition=0, subject=3, eeg_root=str(tmp_path))
    with pytest.raises(RuntimeError, match="Segment length mismatch"):
        seg.extract_2s_segment(block=0, concept=39, repetition=4, data=raw[:, :, :103000])
    np.save(os.path.join(tmp_path, "sub3.npy"), np.arange(7 * 2 * 104000, dtype=np.float32).reshape(7, 2, 104000))
    got = seg.extract_2s_segment(block=2, concept=1, repetition=1, subject=3, eeg_root=str(tmp_path))
    assert got[0, 0] == 2 * 2 * 104000 + 2600 + 600 + 400


def test_seg_sliding_window_view_matches_reference_layout():
    clips = np.arange(7 * 40 * 5 * 62 * 400, dtype=np.float32).reshape(7, 40, 5, 62, 400)
    win = ssw.seg_sliding_window(clips, 0.5, 0.25, fs=200)
    assert win.shape == (7, 40, 5, 7, 62, 100)
    assert win.strides == (19840000, 496000, 99200, 200, 1600, 4)      # SURVEY.md 3.2, verified on the reference
    assert np.shares_memory(win, clips) and not win.flags.writeable
    small = clips[:1, :2]
    assert np.array_equal(ssw.seg_sliding_window(small, 0.5, 0.25), oracle.seg_sliding_window(small, 0.5, 0.25))
    tw = ssw.seg_sliding_window(torch.from_numpy(small), 0.5, 0.25)
    assert tuple(tw.shape) == (1, 2, 5, 7, 62, 100) and np.array_equal(tw.numpy(), oracle.seg_sliding_window(small, 0.5, 0.25))
    with pytest.raises(ValueError):
        ssw.seg_sliding_window(clips[0], 0.5, 0.25)           # needs a 5-D input, like the reference


def test_window_view_is_recognised_and_unwound():
    clips = np.random.default_rng(1).standard_normal((2, 3, 5, 6, 400)).astype(np.float32)
    win = ssw.seg_sliding_window(clips, 0.5, 0.25)
    back = d500._clips_behind_window_view(win)
    assert back is not None and back.shape == clips.shape and np.shares_memory(back, clips)
    assert np.array_equal(back, clips)
    assert d500._clips_behind_window_view(np.ascontiguousarray(win)) is None
    tback = d500._clips_behind_window_view(ssw.seg_sliding_window(torch.from_numpy(clips), 0.5, 0.25))
    assert tback is not None and np.array_equal(tback.numpy(), clips)


def test_shims_validate_before_touching_the_device():
    with pytest.raises(ValueError, match="could not be broadcast"):
        DE_PSD(np.ones((3, 150), np.float32), 200, 1)
    with pytest.raises(ValueError, match="positive"):
        DE_PSD(np.ones((3, 250), np.float32), 0, 1)
    with pytest.raises(ValueError, match="could not be broadcast"):
        DE_PSD(np.ones((3, 50), np.float32), 250, 0.25)                 # int(0.25 * 250) = 62 samples expected
    with pytest.raises(ValueError, match="could not be broadcast"):
        extract_de_psd_raw(np.ones((1, 1, 1, 62, 400), np.float32), fs=250)
    with pytest.raises(ValueError):
        extract_de_psd_raw(np.ones((1, 1, 1, 62, 200), np.float32))
    with pytest.raises(ValueError):
        extract_de_psd_1s(np.ones((1, 1, 62, 400), np.float32))
    with pytest.raises(ValueError, match="could not be broadcast"):
        d500.extract_de_psd_sw(np.ones((1, 1, 1, 7, 62, 90), np.float32), 200, 0.5)


def test_no_silent_cpu_fallback():
    """Without a CUDA device the compute entry points must raise, never compute on the host."""
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DE_PSD(np.ones((3, 200), np.float32), 200, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        extract_de_psd_raw(np.ones((1, 1, 1, 62, 400), np.float32))
    from eeg2video_b200 import ops
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.de_psd_windows(torch.ones(4, 100))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 20, 125, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [cohort.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == cohort.shard_sizes(n, world)


def test_convert_directory_is_the_scripts_file_loop(tmp_path):
    """The shared directory helper behind every script entry point, with a stand-in conversion (no GPU)."""
    import numpy as np
    from eeg2video_b200.EEG_preprocessing import _io
    src = tmp_path / "in"
    src.mkdir()
    for n, shape in ((1, (2, 3)), (2, (2, 3)), (10, (4,))):
        np.save(src / f"sub{n}.npy", np.full(shape, n, dtype=np.float32))
    (src / "notes.txt").write_text("ignored")
    logs = []
    done = _io.convert_directory(str(src), (str(tmp_path / "a"), str(tmp_path / "b")), lambda x: (x + 1, x * 2),
                                 accept=lambda x: None if x.ndim == 2 else f"unexpected shape {x.shape}", log=logs.append)
    assert done == ["sub1.npy", "sub2.npy"]                       # sorted, .npy only, the 1-D file vetoed
    assert any("Skipping sub10.npy: unexpected shape (4,)" in line for line in logs)
    assert np.array_equal(np.load(tmp_path / "a" / "sub2.npy"), np.full((2, 3), 3, np.float32))
    assert np.array_equal(np.load(tmp_path / "b" / "sub1.npy"), np.full((2, 3), 2, np.float32))
    only = _io.convert_directory(str(src), (str(tmp_path / "c"),), lambda x: x, names=["sub2.npy"], log=logs.append)
    assert only == ["sub2.npy"] and not (tmp_path / "c" / "sub1.npy").exists()
    with pytest.raises(FileNotFoundError):
        _io.convert_directory(str(src), (str(tmp_path / "d"),), lambda x: x, names=["sub99.npy"], log=logs.append)


def test_clip_start_matches_the_reference_formula():
    from eeg2video_b200.EEG_preprocessing import segment_raw_signals_200Hz as seg
    for fs in (200, 250, 1000):
        for c in (0, 1, 39):
            for r in (0, 4):
                assert seg.clip_start(c, r, fs) == c * (3 * fs + 5 * 2 * fs) + 3 * fs + r * 2 * fs   # :58-64
    assert seg.clip_start(39, 4) + 400 == 104000


def test_preprocess_all_cli_parses_like_the_scripts():
    from eeg2video_b200 import preprocess_all
    assert set(preprocess_all.FEATURE_DIRS) == {"2s", "1s", "500ms"}
    assert preprocess_all.FEATURE_DIRS["1s"][2].__name__ == "float64"      # the 1 s files are float64 in the reference
    with pytest.raises(SystemExit):
        preprocess_all.main(["--no-such-flag"])


def test_preprocess_all_shards_recordings_by_rank():
    from eeg2video_b200 import preprocess_all
    names = [f"sub{i}.npy" for i in range(1, 11)]
    parts = [preprocess_all.shard_for_this_rank(names, {"WORLD_SIZE": "4", "RANK": str(r)}) for r in range(4)]
    assert sorted(sum(parts, [])) == sorted(names) and [len(p) for p in parts] == [3, 3, 2, 2]
    assert preprocess_all.shard_for_this_rank(names, {}) == names
    with pytest.raises(ValueError):
        preprocess_all.shard_for_this_rank(names, {"WORLD_SIZE": "2", "RANK": "2"})


def test_band_bins_and_hann_follow_the_reference_expressions():
    """The host expressions handed to the general kernel (ops.band_bins, ops.hann_weights) against the oracle's
    restatement of DE_PSD.py:35-39, :51, :63 over a sweep of sampling rates and window lengths."""
    from hypothesis import given, settings, strategies as st
    from eeg2video_b200 import ops

    @settings(max_examples=200, deadline=None)
    @given(st.one_of(st.integers(100, 4000), st.floats(100.0, 4000.0, allow_nan=False)))
    def bins(fre):
        lo, hi = ops.band_bins(fre)
        assert list(zip(lo, hi)) == oracle.band_bin_ranges(fre)
    bins()
    assert ops.band_bins(200) == ([0, 3, 7, 13, 30], [3, 7, 13, 30, 98])
    assert ops.band_bins(250)[0][0] == -1                        # "bin -1": Python's last element, bin 99
    assert max(ops.band_bins(150)[1]) >= 100                     # where the reference raises IndexError
    for length in (1, 7, 100, 199, 200, 201, 640):
        n_live = min(length, 200)
        want = oracle.hann_window(length)[:n_live].astype(np.float32)
        assert np.array_equal(ops.hann_weights(length, n_live), want)
