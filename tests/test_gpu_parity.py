"""Parity of the CUDA path with the oracle and with the reference's golden vectors -- every call goes through the
C ABI of libeegfe.so (via the torch.library ops / the reference-named shims).  Needs a B200: -m gpu.

Bars (BASELINE.json north_star): segmentation bit-exact; PSD <= 1e-4 relative; DE <= 1e-4 absolute (log2).
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import assert_features_close, decode
from eeg2video_b200 import _lib, frontend, ops, synth
from eeg2video_b200.EEG_preprocessing import segment_raw_signals_200Hz as seg
from eeg2video_b200.EEG_preprocessing import segment_sliding_window as ssw
from eeg2video_b200.EEG_preprocessing.DE_PSD import DE_PSD
from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per1s import extract_de_psd_1s
from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per2s import extract_de_psd_raw
from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per500ms import extract_de_psd_sw

pytestmark = pytest.mark.gpu
TW = {100: 0.5, 200: 1, 400: 2}
DEV = "cuda:0"


def test_native_library_is_the_one_in_tree():
    lib = _lib.load()
    assert lib.eegfe_abi_version() == 1
    before = _lib.launch_count()
    ops.de_psd_windows(torch.ones(4, 100, device=DEV) * 3)
    torch.cuda.synchronize()
    assert _lib.launch_count() == before + 1


# ---- DE_PSD itself against the reference's golden outputs ---------------------------------------------------------
@pytest.mark.parametrize("length", (100, 200, 400))
@pytest.mark.parametrize("kind", ("white", "offset", "pink_tone", "small"))
def test_de_psd_golden(golden, length, kind):
    g = golden("de_psd_golden.npz")
    key = f"L{length}_{kind}"
    de, psd = DE_PSD(decode(g[key + "_codes"]), 200, TW[length])
    assert isinstance(de, np.ndarray) and de.dtype == np.float64 and de.shape == (62, 5)
    assert_features_close(de, psd, g[key + "_de"], g[key + "_psd"])


def test_de_psd_impulse_known_answer(golden):
    g = golden("de_psd_golden.npz")
    de, psd = DE_PSD(g["impulse_x"], 200, 0.5)
    assert_features_close(de, psd, g["impulse_de"], g["impulse_psd"])


def test_de_psd_torch_in_torch_out():
    x = torch.randn(62, 200, device=DEV) * 30
    de, psd = DE_PSD(x, 200, 1)
    assert de.is_cuda and de.dtype == torch.float64 and tuple(de.shape) == (62, 5)
    de_ref, psd_ref = oracle.de_psd_closed_form(x.cpu().numpy(), 200, 1)
    assert_features_close(de.cpu().numpy(), psd.cpu().numpy(), de_ref, psd_ref)


@pytest.mark.parametrize("dtype", (np.float64, np.float16, np.int16, np.int32))
def test_de_psd_input_dtypes(dtype):
    rng = np.random.default_rng(3)
    x = (rng.integers(-2000, 2000, (10, 100)) / 8.0).astype(dtype)
    de, psd = DE_PSD(x, 200, 0.5)
    de_ref, psd_ref = oracle.de_psd_closed_form(x.astype(np.float64), 200, 0.5)
    assert_features_close(de, psd, de_ref, psd_ref)


@pytest.mark.parametrize("length", (100, 200, 400))
@pytest.mark.parametrize("n_rows", (1, 2, 31, 33, 63, 64, 65, 127, 129, 1000))
def test_ragged_row_counts(length, n_rows):
    """Tiles hold 32..128 rows: every partial-tile size must be handled."""
    rng = np.random.default_rng(n_rows * 7 + length)
    x = (30 * rng.standard_normal((n_rows, length)) + 5).astype(np.float32)
    de, psd = DE_PSD(x, 200, TW[length])
    de_ref, psd_ref = oracle.de_psd_closed_form(x, 200, TW[length])
    assert_features_close(de, psd, de_ref, psd_ref)


@pytest.mark.parametrize("fre,tw", ((200, 0.25), (200, 0.75), (200, 1.5), (200, 3), (200, 0.005), (250, 1), (250, 0.4),
                                    (256, 0.5), (500, 0.2), (1000, 0.25), (199, 1), (400, 2)))
def test_de_psd_any_window_length_and_rate(fre, tw):
    """DE_PSD accepts any window length and sampling rate (DE_PSD.py:33-39, :49-58): Hann of int(tw * fre) points,
    first 200 samples (zero-padded below), bins from int(f / fre * 200) -- at fre >= 250 the delta band starts at
    "bin -1", Python's last element (bin 99).  General kernel against the loop-for-loop port of the reference."""
    length = int(tw * fre)
    rng = np.random.default_rng(int(fre * 1000 + length))
    x = (30 * rng.standard_normal((9, length)) + rng.uniform(-20, 20, (9, 1))).astype(np.float32)
    de, psd = DE_PSD(x, fre, tw)
    de_ref, psd_ref = oracle.de_psd_loop(x, fre, tw)
    assert de.shape == (9, 5) and de.dtype == np.float64
    assert_features_close(de, psd, de_ref, psd_ref)


def test_de_psd_low_rates_raise_index_error_like_the_reference():
    x = np.ones((2, 150), np.float32)
    with pytest.raises(IndexError, match="out of bounds for axis 0 with size 100"):
        oracle.de_psd_loop(x + np.arange(150, dtype=np.float32), 150, 1)           # the reference's own failure
    with pytest.raises(IndexError, match="out of bounds for axis 0 with size 100"):
        DE_PSD(x, 150, 1)


def test_sliding_driver_other_window_lengths():
    """extract_de_psd_sw(raw, fs, win_sec) for a window the fused kernels do not cover (0.3 s = 60 samples)."""
    rng = np.random.default_rng(5)
    raw = (30 * rng.standard_normal((1, 2, 2, 3, 6, 60))).astype(np.float32)
    de, psd = extract_de_psd_sw(raw, 200, 0.3)
    de_ref, psd_ref = oracle.extract_de_psd_sw(raw, 200, 0.3, closed=False)
    assert de.dtype == np.float32 and de.shape == (1, 2, 2, 3, 6, 5)
    assert_features_close(de, psd, de_ref, psd_ref)


def test_zero_power_raises_like_the_reference():
    x = np.ones((4, 200), np.float32)
    x[2] = 0
    with pytest.raises(ValueError, match="math domain error"):
        DE_PSD(x, 200, 1)
    with pytest.raises(ValueError, match="math domain error"):
        oracle.de_psd_loop(x, 200, 1)


def test_empty_inputs():
    de, psd = DE_PSD(np.zeros((0, 100), np.float32), 200, 0.5)
    assert de.shape == (0, 5) and psd.shape == (0, 5)
    de, psd, _ = ops.de_psd_from_clips(torch.zeros(0, 62, 400, device=DEV), _lib.MODE_500MS)
    assert tuple(de.shape) == (0, 7, 62, 5)


def test_strided_rows_and_unaligned_rows():
    """Row stride != length (a view into a wider buffer) and rows that are only 4-byte aligned (odd stride /
    offset base): the first uses TMA bulk copies, the second the cooperative-load path; same numbers."""
    rng = np.random.default_rng(11)
    wide = torch.from_numpy((30 * rng.standard_normal((70, 404))).astype(np.float32)).to(DEV)
    ref = oracle.de_psd_closed_form(wide[:, :100].cpu().numpy(), 200, 0.5)
    de, psd, _ = ops.de_psd_windows(wide[:, :100])
    assert_features_close(de.cpu().numpy(), psd.cpu().numpy(), *ref)
    odd = torch.from_numpy((30 * rng.standard_normal((70, 203))).astype(np.float32)).to(DEV)
    view = odd[:, 1:201]
    ref = oracle.de_psd_closed_form(view.cpu().numpy(), 200, 1)
    de, psd, _ = ops.de_psd_windows(view)
    assert_features_close(de.cpu().numpy(), psd.cpu().numpy(), *ref)
    aligned_copy = view.contiguous()
    de2, psd2, _ = ops.de_psd_windows(aligned_copy)
    assert torch.equal(de, de2) and torch.equal(psd, psd2)       # both load paths feed identical arithmetic


# ---- drivers against the reference's golden outputs -----------------------------------------------------------------
def test_drivers_golden(golden):
    g = golden("drivers_golden.npz")
    clips = decode(g["codes"])
    de, psd = extract_de_psd_raw(clips, 200)
    assert de.dtype == np.float32 and de.shape == (2, 2, 3, 62, 5)
    assert_features_close(de, psd, g["de_2s"], g["psd_2s"])
    de, psd = extract_de_psd_1s(clips, 200)
    assert de.dtype == np.float64 and de.shape == (2, 2, 3, 2, 62, 5)
    assert_features_close(de, psd, g["de_1s"], g["psd_1s"])
    win_view = ssw.seg_sliding_window(clips, 0.5, 0.25, fs=200)
    de, psd = extract_de_psd_sw(win_view, 200, 0.5)                      # strided view -> fused path
    assert de.dtype == np.float32 and de.shape == (2, 2, 3, 7, 62, 5)
    assert_features_close(de, psd, g["de_500ms"], g["psd_500ms"])
    de_m, psd_m = extract_de_psd_sw(np.ascontiguousarray(win_view), 200, 0.5)   # materialised -> pre-cut path
    assert np.array_equal(de, de_m) and np.array_equal(psd, psd_m)


def test_one_second_script_golden(golden):
    g = golden("script_1s_golden.npz")
    clips = decode(g["clips_codes"])[None, None]
    de, psd = extract_de_psd_1s(clips, 200)
    assert_features_close(de[0, 0], psd[0, 0], g["de"], g["psd"])


# ---- segmentation: bit-exact -------------------------------------------------------------------------------------------
def _formula_raw(n_ch, n_t):
    b = np.arange(7).reshape(7, 1, 1)
    ch = np.arange(n_ch).reshape(1, n_ch, 1)
    t = np.arange(n_t).reshape(1, 1, n_t)
    return (((b * 7919 + ch * 104729 + t * 31) % 65536) - 32768).astype(np.int16)


def test_segment_clips_golden_bit_exact(golden):
    g = golden("segment_golden.npz")
    raw = _formula_raw(int(g["n_ch"]), int(g["n_t"]))
    clips = seg.segment_subject(raw)
    assert clips.dtype == np.int16 and clips.shape == (7, 40, 5, int(g["n_ch"]), 400)
    for (b, c, r), want in zip(g["picks"], g["segments"]):
        assert np.array_equal(clips[b, c, r], want)
    assert np.array_equal(clips, oracle.segment_subject(raw))


@pytest.mark.parametrize("dtype", (np.float32, np.float64, np.float16, np.int16))
@pytest.mark.parametrize("n_t", (104000, 104001, 104003, 110000))
def test_segment_clips_dtypes_and_odd_lengths(dtype, n_t):
    rng = np.random.default_rng(n_t)
    raw = rng.integers(-30000, 30000, (7, 3, n_t)).astype(dtype)
    assert np.array_equal(seg.segment_subject(raw), oracle.segment_subject(raw))


def test_segment_all_files_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    raw = rng.standard_normal((7, 62, 104000)).astype(np.float32)
    src, dst = tmp_path / "EEG", tmp_path / "out"
    src.mkdir()
    np.save(src / "sub7.npy", raw)
    seg.segment_all_files(str(src), str(dst), 200)
    got = np.load(dst / "sub7.npy")
    assert got.dtype == np.float32 and np.array_equal(got, oracle.segment_subject(raw))


def test_block_too_short():
    raw = torch.zeros(1, 62, 103999, device=DEV)
    with pytest.raises(RuntimeError, match="Segment length mismatch"):
        frontend.de_psd_from_raw(raw, "500ms")
    with pytest.raises(RuntimeError, match="Segment length mismatch"):
        ops.segment_clips(raw, 200)
    with pytest.raises(RuntimeError, match="Segment length mismatch"):
        seg.segment_subject(np.zeros((7, 2, 1000), np.float32))


def test_sliding_windows_materialised_bit_exact():
    rng = np.random.default_rng(9)
    for dtype in (np.float32, np.float64, np.int16):
        clips = rng.integers(-9999, 9999, (2, 3, 5, 62, 400)).astype(dtype)
        got = ssw.materialize_windows(clips)
        assert np.array_equal(got, oracle.seg_sliding_window(clips, 0.5, 0.25))


# ---- the fused path from raw recordings --------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", (np.float32, np.float64, np.int16))
def test_seq2seq_window_layout_bit_exact(golden, dtype):
    """(.., ch, 400) -> (.., ch, 100, 7), the Seq2Seq trainer's layout (my_autoregressive_transformer.py:309-314):
    bit-exact against the golden made by the reference's own lines and against the oracle on a bigger tensor."""
    g = golden("seq2seq_golden.npz")
    got = frontend.sliding_windows(torch.from_numpy(g["codes"].astype(dtype)).to(DEV), layout="last")
    assert tuple(got.shape) == (2, 3, 5, 100, 7)
    assert np.array_equal(got.cpu().numpy(), g["windows"].astype(dtype))
    rng = np.random.default_rng(11)
    clips = rng.integers(-30000, 30000, (3, 7, 5, 62, 400)).astype(dtype)
    got = frontend.sliding_windows(torch.from_numpy(clips).to(DEV), layout="last").cpu().numpy()
    assert np.array_equal(got, oracle.seq2seq_windows(clips))
    # same bytes as the window-major form, axes permuted
    major = frontend.sliding_windows(torch.from_numpy(clips).to(DEV)).cpu().numpy()          # (.., 7, 62, 100)
    assert np.array_equal(np.moveaxis(major, -3, -1), got)


@pytest.fixture(scope="module")
def subject():
    raw = synth.synth_subject(1, device=DEV)
    return raw, raw.cpu().numpy()


@pytest.mark.parametrize("mode,win,hop,nwin", (("500ms", 100, 50, 7), ("1s", 200, 200, 2), ("2s", 400, 0, 1)))
def test_fused_from_raw_full_subject(subject, mode, win, hop, nwin):
    """One full-size subject (7 x 62 x 104000): every one of the 86 800 * nwin channel-windows against the
    float64 closed form of the reference arithmetic."""
    raw, raw_np = subject
    de, psd = frontend.de_psd_from_raw(raw, mode)
    want_shape = (7, 40, 5) + ((nwin,) if nwin > 1 else ()) + (62, 5)
    assert tuple(de.shape) == want_shape and de.dtype == torch.float32
    clips = oracle.segment_subject(raw_np)                                         # (7,40,5,62,400)
    if nwin > 1:
        wins = np.stack([clips[..., w * hop:w * hop + win] for w in range(nwin)], axis=3)   # (7,40,5,W,62,win)
    else:
        wins = clips
    de_ref, psd_ref = oracle.de_psd_closed_form(wins, 200, TW[win])
    assert_features_close(de.cpu().numpy(), psd.cpu().numpy(), de_ref, psd_ref)


def test_fused_equals_staged_bit_for_bit(subject):
    """raw -> features  ==  raw -> segment_clips -> features  ==  raw -> clips -> sliding windows -> features."""
    raw, _ = subject
    blocks = raw[:2]
    de, psd = frontend.de_psd_from_raw(blocks, "500ms")
    clips = frontend.segment_clips(blocks)
    de_c, psd_c = frontend.de_psd_from_clips(clips, "500ms")
    assert torch.equal(de, de_c) and torch.equal(psd, psd_c)
    wins = frontend.sliding_windows(clips)
    de_w, psd_w = frontend.de_psd_windows(wins)
    assert torch.equal(de, de_w) and torch.equal(psd, psd_w)
    for mode in ("1s", "2s"):
        a = frontend.de_psd_from_raw(blocks, mode)
        b = frontend.de_psd_from_clips(clips, mode)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_fused_against_loop_oracle_sample(subject):
    """A random sample of clips against the loop-for-loop port of the reference (scipy.fftpack, math.log)."""
    raw, raw_np = subject
    de, psd = frontend.de_psd_from_raw(raw, "500ms")
    de, psd = de.cpu().numpy(), psd.cpu().numpy()
    rng = np.random.default_rng(0)
    for _ in range(6):
        b, c, r, w = rng.integers(7), rng.integers(40), rng.integers(5), rng.integers(7)
        clip = oracle.extract_2s_segment(block=int(b), concept=int(c), repetition=int(r), data=raw_np)
        de_ref, psd_ref = oracle.de_psd_loop(clip[:, 50 * w:50 * w + 100], 200, 0.5)
        assert_features_close(de[b, c, r, w], psd[b, c, r, w], de_ref, psd_ref)


def test_properties_at_full_size(subject):
    """Size-independent properties on a full subject: exact x4 under x2 scaling, channel-permutation
    equivariance, independence of the 2 s features from the clip's second half, determinism."""
    raw, _ = subject
    de, psd = frontend.de_psd_from_raw(raw, "500ms")
    de2, psd2 = frontend.de_psd_from_raw(raw * 2, "500ms")
    assert torch.equal(psd2, psd * 4)
    assert torch.allclose(de2, de + 2, atol=2e-6)
    perm = torch.randperm(62, device=DEV)
    de_p, psd_p = frontend.de_psd_from_raw(raw[:, perm], "500ms")
    assert torch.equal(psd_p, psd[..., perm, :]) and torch.equal(de_p, de[..., perm, :])
    again = frontend.de_psd_from_raw(raw, "500ms")
    assert torch.equal(again[0], de) and torch.equal(again[1], psd)
    clips = frontend.segment_clips(raw[:1])
    a = frontend.de_psd_from_clips(clips, "2s")
    clips[..., 200:] = 1e6
    b = frontend.de_psd_from_clips(clips, "2s")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_leading_axes_and_odd_channel_counts():
    raw = synth.synth_blocks(3, 77, device=DEV, channels=5)
    de, psd = frontend.de_psd_from_raw(raw.reshape(1, 3, 5, -1), "1s")
    assert tuple(de.shape) == (1, 3, 40, 5, 2, 5, 5)
    clips = oracle.segment_subject(np.concatenate([raw.cpu().numpy()] + [np.zeros((4, 5, 104000), np.float32)]))[:3]
    wins = np.stack([clips[..., :200], clips[..., 200:]], axis=3)
    de_ref, psd_ref = oracle.de_psd_closed_form(wins, 200, 1)
    assert_features_close(de[0].cpu().numpy(), psd[0].cpu().numpy(), de_ref, psd_ref)


@pytest.mark.parametrize("n_clips,n_ch", ((1, 1), (1, 3), (2, 5), (3, 16), (5, 17), (7, 31), (1, 62), (9, 62), (50, 64),
                                          (33, 7), (200, 2)))
def test_streaming_kernel_ragged_tiles(n_clips, n_ch):
    """500 ms streaming kernel (16-row tiles, 7 half-passes per tile, passes straddling tiles): row counts that
    leave the last tile partial, odd numbers of half-passes per CTA, tiles spanning several clips (n_ch < 16)
    and tiles cut by a clip boundary -- every channel-window against the float64 closed form."""
    rng = np.random.default_rng(1000 * n_clips + n_ch)
    clips = (30 * rng.standard_normal((n_clips, n_ch, 400)) + rng.uniform(-50, 50, (n_clips, n_ch, 1))).astype(np.float32)
    de, psd = frontend.de_psd_from_clips(torch.from_numpy(clips).to(DEV), "500ms")
    assert tuple(de.shape) == (n_clips, 7, n_ch, 5)
    wins = np.stack([clips[..., 50 * w:50 * w + 100] for w in range(7)], axis=1)        # (n, 7, ch, 100)
    de_ref, psd_ref = oracle.de_psd_closed_form(wins, 200, 0.5)
    assert_features_close(de.cpu().numpy(), psd.cpu().numpy(), de_ref, psd_ref)


def test_streaming_kernel_many_tiles_per_cta():
    """More tiles than ring slots on every SM (2 subjects = 10850 tiles / 148 SMs = 73 per CTA, 10 generations of
    the 7-slot ring), twice in a row on the same buffers: slot recycling, staging drain flags, determinism."""
    raw = synth.synth_blocks(14, 9, device=DEV)
    a = frontend.de_psd_from_raw(raw, "500ms")
    b = frontend.de_psd_from_raw(raw, "500ms")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    clips = frontend.segment_clips(raw[3])
    wins = frontend.sliding_windows(clips)                                              # (40,5,7,62,100)
    de_w, psd_w = frontend.de_psd_windows(wins)                                         # ring kernel, pre-cut windows
    assert torch.equal(a[0][3], de_w) and torch.equal(a[1][3], psd_w)


@pytest.fixture
def tensor_loads():
    old = _lib.set_tensor_loads(True)
    yield
    _lib.set_tensor_loads(old)


def test_tensor_copy_producer_serves_the_200_sample_rows(subject, tensor_loads):
    """Measurement option eegfe_set_tensor_loads(1): 2 s mode and pre-cut 200 / 400-sample windows fetch a tile with
    ONE TMA tensor copy (clip-aligned tiles, csrc/eegfe_kernels.cu `attach_tensor_map`); jobs with fewer than 24
    channels stay on per-row bulk copies.  Same arithmetic either way: bit-identical, channel for channel."""
    raw, _ = subject
    _lib.set_tensor_loads(False)
    before = _lib.tma_launch_count()
    de0, psd0 = frontend.de_psd_from_raw(raw[:2], "2s")                              # the default: per-row bulk copies
    assert _lib.tma_launch_count() == before
    _lib.set_tensor_loads(True)
    de, psd = frontend.de_psd_from_raw(raw[:2], "2s")
    assert _lib.tma_launch_count() == before + 1
    assert torch.equal(de, de0) and torch.equal(psd, psd0)
    few = raw[:2, :20].contiguous()                                                  # 20 channels -> bulk-copy producer
    before = _lib.tma_launch_count()
    de_f, psd_f = frontend.de_psd_from_raw(few, "2s")
    assert _lib.tma_launch_count() == before
    assert torch.equal(de_f, de[..., :20, :]) and torch.equal(psd_f, psd[..., :20, :])
    # pre-cut windows: uniformly strided rows -> tensor boxes; a pitch that is not a multiple of 16 bytes -> fallback
    clips = frontend.segment_clips(raw[:1])
    before = _lib.tma_launch_count()
    a = ops.de_psd_windows(clips.reshape(-1, 400))
    assert _lib.tma_launch_count() == before + 1
    assert torch.equal(a[0].reshape(de[0].shape), de[0]) and torch.equal(a[1].reshape(de[0].shape), psd[0])
    padded = torch.empty((clips.numel() // 400, 402), device=DEV)
    padded[:, :400] = clips.reshape(-1, 400)
    b = ops.de_psd_windows(padded[:, :400])
    assert _lib.tma_launch_count() == before + 1
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    one = ops.de_psd_windows(clips.reshape(-1, 200))                                 # 1 s windows, dense rows
    de1, psd1 = frontend.de_psd_from_raw(raw[:1], "1s")
    shape = (1, 40, 5, 62, 2, 5)                                                     # rows run (clip, channel, half)
    assert torch.equal(one[0].reshape(shape).permute(0, 1, 2, 4, 3, 5), de1)
    assert torch.equal(one[1].reshape(shape).permute(0, 1, 2, 4, 3, 5), psd1)


@pytest.mark.parametrize("n_clips,n_ch", ((1, 24), (3, 31), (147, 32), (149, 33), (300, 48), (1000, 62), (700, 64),
                                          (31, 65), (40, 100)))
def test_clip_aligned_tiles_ragged(n_clips, n_ch, tensor_loads):
    """2 s ring kernel with tensor boxes: tiles are 32 channels of ONE clip, the last tile of a clip is short (62 = 32
    + 30, 65 = 32 + 32 + 1); fewer tiles than CTAs, tile counts that do not divide by the grid, slot recycling."""
    rng = np.random.default_rng(77 * n_clips + n_ch)
    clips = (30 * rng.standard_normal((n_clips, n_ch, 400)) + rng.uniform(-50, 50, (n_clips, n_ch, 1))).astype(np.float32)
    before = _lib.tma_launch_count()
    de, psd = frontend.de_psd_from_clips(torch.from_numpy(clips).to(DEV), "2s")
    assert _lib.tma_launch_count() == before + 1
    de_ref, psd_ref = oracle.de_psd_closed_form(clips, 200, 2)
    assert_features_close(de.cpu().numpy(), psd.cpu().numpy(), de_ref, psd_ref)


def test_unaligned_block_length_uses_fallback_loader():
    """T = 104001: rows are only 4-byte aligned, so TMA bulk copies are impossible; same results required."""
    raw = synth.synth_blocks(1, 5, device=DEV, channels=62, block_len=104001)
    de, psd = frontend.de_psd_from_raw(raw, "500ms")
    aligned = raw[..., :104000].contiguous()
    de_a, psd_a = frontend.de_psd_from_raw(aligned, "500ms")
    assert torch.equal(de, de_a) and torch.equal(psd, psd_a)


@pytest.mark.parametrize("mode", ("500ms", "1s", "2s"))
@pytest.mark.parametrize("t_len,offset", ((104001, 0), (104002, 0), (104003, 0), (104003, 1), (104002, 2), (104001, 3)))
def test_rows_of_every_alignment(mode, t_len, offset):
    """Rows that start 4, 8 or 12 bytes off a 16-byte boundary -- an odd block length (the offset then changes from
    row to row) or a view that starts inside a buffer of pitch 104004 (the same offset for every row).  TMA cannot start a copy there:
    the kernels copy the 16-byte aligned span around each row and read it shifted.  Called through the C ABI directly
    and through the op; both must equal the result on an aligned copy, bit for bit.  37 channels: ragged tiles."""
    base = synth.synth_blocks(1, 9, device=DEV, channels=37, block_len=t_len + offset)
    raw = base[..., offset:]
    assert raw.data_ptr() % 16 == 4 * offset
    aligned = raw[..., :104000].contiguous()
    want = ops.de_psd_from_raw(aligned, frontend.MODES[mode])
    got = ops.de_psd_from_raw(raw, frontend.MODES[mode])
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    de, psd = torch.zeros_like(want[0]), torch.zeros_like(want[1])
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    before = _lib.launch_count()
    _lib.check(_lib.load().eegfe_de_psd_from_raw(raw.data_ptr(), raw.shape[0], raw.shape[1], raw.shape[2], raw.stride(0),
                                                 raw.stride(1), frontend.MODES[mode], de.data_ptr(), psd.data_ptr(),
                                                 status.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert _lib.launch_count() == before + 1                                   # one kernel, no re-alignment pass
    assert torch.equal(de, want[0]) and torch.equal(psd, want[1]) and int(status.item()) == 0


def test_shifted_rows_at_the_end_of_an_allocation():
    """The aligned span around the LAST row of a buffer ends at most 12 bytes past the row -- inside the allocation,
    whose size is a multiple of 16 bytes: a recording that fills its allocation exactly, last rows included."""
    n_ch, t_len = 62, 104001
    flat = torch.empty(2 * n_ch * t_len, device=DEV)                           # torch rounds the block up, not down
    raw = flat.view(2, n_ch, t_len)
    raw.copy_(synth.synth_blocks(2, 13, device=DEV, channels=n_ch, block_len=t_len))
    for mode in ("1s", "2s"):
        want = ops.de_psd_from_raw(raw[..., :104000].contiguous(), frontend.MODES[mode])
        got = ops.de_psd_from_raw(raw, frontend.MODES[mode])
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


# ---- host pipeline / compact layout -----------------------------------------------------------------------------------
def test_from_concepts_equals_from_raw(subject):
    raw, _ = subject
    blocks = raw[:3]
    compact = blocks.reshape(3, 62, 40, 2600)[..., 600:].contiguous()          # (3, 62, 40, 2000)
    for mode in ("500ms", "1s", "2s"):
        a = ops.de_psd_from_raw(blocks, frontend.MODES[mode])
        b = ops.de_psd_from_concepts(compact, frontend.MODES[mode])
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("compact", (True, False, "auto"))
@pytest.mark.parametrize("block_len", (104000, 104400))
def test_host_pipeline_matches_device_path(compact, block_len):
    from eeg2video_b200 import pipeline
    raw = synth.synth_blocks(5, 31, device=DEV, block_len=block_len)
    want_de, want_psd = frontend.de_psd_from_raw(raw, "500ms")
    de, psd = pipeline.features_from_host(raw.cpu(), "500ms", chunk_blocks=2, device=DEV, compact=compact)
    assert not de.is_cuda and tuple(de.shape) == (5, 40, 5, 7, 62, 5)
    assert torch.equal(de, want_de.cpu()) and torch.equal(psd, want_psd.cpu())


def test_host_pipeline_picks_its_upload_layout_once():
    """compact="auto": the first run() times the strided and the contiguous upload of its first chunk and keeps the
    faster; later pipelines of the same geometry on the same device reuse the decision (no second probe)."""
    from eeg2video_b200 import pipeline
    pipeline.HostPipeline._layout_cache.clear()
    raw = synth.synth_blocks(3, 32, device="cpu").pin_memory()
    pipe = pipeline.HostPipeline(DEV, 62, 104000, chunk_blocks=2, mode="1s")
    assert pipe.compact is None and pipe.upload_probe is None
    de = torch.empty(pipe.feature_shape(3)).pin_memory()
    psd = torch.empty_like(de).pin_memory()
    assert pipe.run(raw, de, psd) == 0
    probe = pipe.upload_probe
    assert pipe.compact in (True, False) and probe["blocks"] == 2
    assert pipe.compact == (probe["strided_live_samples_ms"] <= probe["contiguous_rows_ms"])
    assert pipe.h2d_bytes(3) == 3 * 62 * (80000 if pipe.compact else 104000) * 4
    want = frontend.de_psd_from_raw(raw.to(DEV), "1s")
    assert torch.equal(de, want[0].cpu().reshape(de.shape)) and torch.equal(psd, want[1].cpu().reshape(psd.shape))
    again = pipeline.HostPipeline(DEV, 62, 104000, chunk_blocks=2, mode="2s")
    assert again.compact == pipe.compact and again.upload_probe is probe


def test_host_pipeline_zero_power_flag():
    from eeg2video_b200 import pipeline
    raw = synth.synth_blocks(1, 3, device="cpu")
    raw[0, 5, 600:1000] = 0.0
    with pytest.raises(ValueError, match="math domain error"):
        pipeline.features_from_host(raw, "2s", device=DEV)


# ---- next row: GLMNet input build ---------------------------------------------------------------------------------------
def test_glmnet_channel_stats_and_inputs(subject):
    from eeg2video_b200 import glmnet_inputs
    from oracle import glmnet_inputs as oracle_glm
    raw, raw_np = subject
    train = [0, 1, 2, 3, 4, 6]                                # leave block 5 out
    mean, std = glmnet_inputs.channel_stats(raw, train)
    mean_ref, std_ref = oracle_glm.channel_stats(raw_np, train)
    assert np.allclose(mean.cpu().numpy(), mean_ref, rtol=0, atol=1e-9 * np.abs(mean_ref).max() + 1e-9)
    assert np.allclose(std.cpu().numpy(), std_ref, rtol=1e-10)
    mask = torch.zeros(7, dtype=torch.bool, device=DEV)
    mask[train] = True
    mean2, std2 = glmnet_inputs.channel_stats(raw, mask)
    assert torch.equal(mean, mean2) and torch.equal(std, std2)

    clips, de, psd = glmnet_inputs.build_inputs(raw, mean, std)
    assert tuple(clips.shape) == (7, 40, 5, 1, 62, 400) and tuple(de.shape) == (7, 40, 5, 7, 62, 5)
    want = oracle_glm.normalised_clips(raw_np, mean_ref, std_ref)
    assert np.max(np.abs(clips.cpu().numpy() - want)) <= 2e-6           # fp32 fma(x, 1/std, -mean/std) vs float64
    de0, psd0 = frontend.de_psd_from_raw(raw, "500ms")
    assert torch.equal(de, de0) and torch.equal(psd, psd0)              # the feature product is unchanged
    # the normalised clips have zero mean / unit std per channel over the training blocks
    x = clips[train].double()
    assert x.mean(dim=(0, 1, 2, 3, 5)).abs().max() < 1e-5
    assert (x.std(dim=(0, 1, 2, 3, 5), unbiased=False) - 1).abs().max() < 1e-5


def test_glmnet_inputs_large_offset_and_constant_channel():
    """(x - mean) * (1 / std) subtracts FIRST: a channel riding on a DC offset 1000x its std keeps its digits; a constant
    channel (std == 0) comes out as zeros, not inf / NaN (scale 1, scikit-learn's convention)."""
    from eeg2video_b200 import glmnet_inputs
    from oracle import glmnet_inputs as oracle_glm
    raw = synth.synth_blocks(1, 31, device=DEV, channels=16)
    raw[:, 3] = raw[:, 3] * 0.1 + 30000.0            # std 3, offset 30000
    raw[:, 5] = 7.5                                  # constant channel
    mean, std = glmnet_inputs.channel_stats(raw)
    assert float(std[5]) == 0.0
    clips, _, _ = glmnet_inputs.build_inputs(raw, mean, std, check=False)
    got = clips.cpu().numpy()
    assert np.all(np.isfinite(got)) and np.all(got[..., 5, :] == 0.0)
    std_np = std.cpu().numpy().copy()
    std_np[5] = 1.0
    want = oracle_glm.normalised_clips(raw.cpu().numpy(), mean.cpu().numpy(), std_np)
    assert np.max(np.abs(got - want)) <= 2e-3 * 1.0 and np.max(np.abs(got[..., 3, :] - want[..., 3, :])) <= 1e-3
    keep = [c for c in range(16) if c not in (3, 5)]
    assert np.max(np.abs(got[..., keep, :] - want[..., keep, :])) <= 4e-6


def test_glmnet_inputs_need_aligned_rows():
    from eeg2video_b200 import glmnet_inputs
    raw = synth.synth_blocks(1, 3, device=DEV, channels=4, block_len=104001)
    with pytest.raises(RuntimeError, match="invalid argument"):
        glmnet_inputs.build_inputs(raw, torch.zeros(4), torch.ones(4))


# ---- script entry points (A8): directory conventions and CLI flags of the reference ------------------------------------
def test_script_entry_points_follow_the_reference_layout(tmp_path, monkeypatch):
    """segment_all_files -> sliding-window script -> the three feature scripts, all on ./data/... defaults."""
    from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per1s as s1
    from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per2s as s2
    from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per500ms as s5
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(17)
    (tmp_path / "data" / "EEG").mkdir(parents=True)
    for sub in (1, 2):
        np.save(tmp_path / "data" / "EEG" / f"sub{sub}.npy", (30 * rng.standard_normal((7, 3, 104000))).astype(np.float32))
    seg.segment_all_files()                                                       # ./data/EEG -> Segmented_Rawf_200Hz_2s
    pre = tmp_path / "data" / "Preprocessing"
    clips = np.load(pre / "Segmented_Rawf_200Hz_2s" / "sub2.npy")
    assert clips.shape == (7, 40, 5, 3, 400) and clips.dtype == np.float32
    assert ssw.main() == ["sub1.npy", "sub2.npy"]                                 # -> Segmented_500ms_sw
    wins = np.load(pre / "Segmented_500ms_sw" / "sub2.npy")
    assert wins.shape == (7, 40, 5, 7, 3, 100) and np.array_equal(wins, oracle.seg_sliding_window(clips, 0.5, 0.25))
    assert s5.main(["--subs", "2"]) == ["sub2.npy"]                               # CLI flags of the reference
    assert s2.main(subjects=(1, 2)) == ["sub1.npy", "sub2.npy"]
    assert s1.main() == ["sub1.npy", "sub2.npy"]
    de5 = np.load(pre / "DE_500ms_sw" / "sub2.npy")
    de2 = np.load(pre / "DE_1per2s" / "sub2.npy")
    psd1 = np.load(pre / "PSD_1per1s" / "sub2.npy")
    assert de5.shape == (7, 40, 5, 7, 3, 5) and de5.dtype == np.float32
    assert de2.shape == (7, 40, 5, 3, 5) and de2.dtype == np.float32
    assert psd1.shape == (7, 40, 5, 2, 3, 5) and psd1.dtype == np.float64
    assert not (pre / "DE_500ms_sw" / "sub1.npy").exists()                        # --subs 2 only
    de_ref, _ = oracle.de_psd_closed_form(wins[0, 0], 200, 0.5)
    assert np.max(np.abs(de5[0, 0] - de_ref)) <= 1e-4


@pytest.mark.parametrize("mode", ("500ms", "1s", "2s"))
def test_repeated_launches_are_bit_identical(mode):
    """Race canary: 40 launches on batches of varying size (different tiles per CTA, different tails) must all
    reproduce the same bits.  (A fifth worker group in the 1 s ring kernel once corrupted one tile in ~7 % of the
    launches: worker groups waited on an mbarrier parity two phases ahead -- tools/stress.py is the long form.)"""
    raw = synth.synth_cohort(range(4), DEV).reshape(28, 62, 104000)
    mid = frontend.MODES[mode]
    cands = [ops.de_psd_from_raw(raw, mid) for _ in range(3)]
    ref = cands[0] if torch.equal(cands[0][0], cands[1][0]) or torch.equal(cands[0][0], cands[2][0]) else cands[1]
    for i in range(40):
        n = 28 - (i % 5)
        de, psd, _ = ops.de_psd_from_raw(raw[:n], mid)
        assert torch.equal(de, ref[0][:n * 200]) and torch.equal(psd, ref[1][:n * 200]), f"launch {i}"


def test_float64_and_int16_recordings():
    """The reference takes whatever dtype the .npy holds; the front end rounds to float32 on the device."""
    from eeg2video_b200 import pipeline
    rng = np.random.default_rng(23)
    raw64 = (30 * rng.standard_normal((2, 5, 104000))).astype(np.float64)
    want = frontend.de_psd_from_raw(torch.from_numpy(raw64.astype(np.float32)).to(DEV), "1s")
    de, psd = pipeline.features_from_host(raw64, "1s", device=DEV)
    assert torch.equal(de, want[0].cpu()) and torch.equal(psd, want[1].cpu())
    clips64 = oracle.segment_subject(np.concatenate([raw64, np.zeros((5, 5, 104000))]))[:2]
    de2, psd2 = extract_de_psd_1s(clips64, 200)
    assert de2.dtype == np.float64 and np.array_equal(de2.astype(np.float32), want[0].cpu().numpy())
    codes = rng.integers(-3000, 3000, (40, 100)).astype(np.int16)
    a = DE_PSD(codes, 200, 0.5)
    b = DE_PSD(codes.astype(np.float64), 200, 0.5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_nan_samples_stay_local_like_the_reference():
    """A NaN sample poisons exactly the channel-windows that contain it (the reference: NaN through the FFT, math.log(nan)
    = nan, no exception); every other channel-window is unchanged."""
    rng = np.random.default_rng(29)
    clips = (30 * rng.standard_normal((3, 4, 400))).astype(np.float32)
    clean = frontend.de_psd_from_clips(torch.from_numpy(clips).to(DEV), "500ms")
    clips[1, 2, 120] = np.nan                                       # inside windows 1 and 2 (50..149, 100..199) of clip 1, ch 2
    de, psd = frontend.de_psd_from_clips(torch.from_numpy(clips).to(DEV), "500ms")
    bad = torch.isnan(de).any(dim=-1)
    want = torch.zeros_like(bad)
    want[1, 1, 2] = want[1, 2, 2] = True
    assert torch.equal(bad, want) and torch.equal(torch.isnan(psd).any(dim=-1), want)
    assert torch.equal(de[~bad], clean[0][~bad])
    de_ref, psd_ref = oracle.de_psd_loop(clips[1, :, 50:150], 200, 0.5)
    assert np.isnan(de_ref[2]).all() and not np.isnan(de_ref[[0, 1, 3]]).any()


def test_jobs_larger_than_32_bit_indexing():
    """5000 blocks = 1 000 000 clips = 2.17e9 feature values per array: the launcher must split the job (row and
    output indices are 32-bit inside the kernels).  All blocks alias one recording (block stride 0), so every block's
    features must equal block 0's."""
    free, _ = torch.cuda.mem_get_info()
    if free < 24e9:
        pytest.skip("needs ~18 GB of free device memory")
    one = synth.synth_blocks(1, 41, device=DEV)                       # (1, 62, 104000)
    de, psd, status = ops.de_psd_from_raw(one.expand(5000, 62, 104000), _lib.MODE_500MS)
    assert de.shape[0] == 1_000_000 and de.numel() > 2 ** 31
    ref_de, ref_psd, _ = ops.de_psd_from_raw(one, _lib.MODE_500MS)
    de = de.reshape(5000, 200, 7, 62, 5)
    psd = psd.reshape(5000, 200, 7, 62, 5)
    for b in (0, 1, 4947, 4948, 4949, 4999):                         # 4948 blocks fit one launch; the seam is at 4948
        assert torch.equal(de[b], ref_de) and torch.equal(psd[b], ref_psd), b
    assert bool((de == ref_de.unsqueeze(0)).all())
    assert int(status.item()) == 0
    del de, psd
    torch.cuda.empty_cache()


def test_launches_are_cuda_graph_capturable():
    """The C-ABI launchers only enqueue work on the given stream (no allocation, no synchronisation), so a whole
    feature pass can be captured once and replayed: new recordings are copied into the captured input buffer."""
    raw = synth.synth_blocks(3, 51, device=DEV)
    other = synth.synth_blocks(3, 52, device=DEV)
    static_in = raw.clone()
    for mode in ("500ms", "1s"):
        mid = frontend.MODES[mode]
        ops.de_psd_from_raw(static_in, mid)                              # first call configures the kernel attributes
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            de, psd, status = ops.de_psd_from_raw(static_in, mid)
        static_in.copy_(other)
        graph.replay()
        torch.cuda.synchronize()
        want = ops.de_psd_from_raw(other, mid)
        assert torch.equal(de, want[0]) and torch.equal(psd, want[1]) and int(status.item()) == 0
        static_in.copy_(raw)
        graph.replay()
        want = ops.de_psd_from_raw(raw, mid)
        assert torch.equal(de, want[0]) and torch.equal(psd, want[1])


def test_preprocess_all_equals_the_five_scripts(tmp_path, monkeypatch):
    """One pass from raw recordings == the reference's chain of scripts, file for file."""
    from eeg2video_b200 import preprocess_all
    from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per1s as s1
    from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per2s as s2
    from eeg2video_b200.EEG_preprocessing import extract_DE_PSD_features_1per500ms as s5
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(19)
    (tmp_path / "data" / "EEG").mkdir(parents=True)
    np.save(tmp_path / "data" / "EEG" / "sub4.npy", (30 * rng.standard_normal((7, 3, 104004))).astype(np.float64))
    assert preprocess_all.main(["--out_root", str(tmp_path / "fused"), "--keep-segments"]) == ["sub4.npy"]
    seg.segment_all_files()
    ssw.main()
    s5.main(["--subs", "4"])
    s2.main(subjects=(4,))
    s1.main()
    chain = tmp_path / "data" / "Preprocessing"
    for d in ("DE_1per2s", "PSD_1per2s", "DE_1per1s", "PSD_1per1s", "DE_500ms_sw", "PSD_500ms_sw",
              "Segmented_500ms_sw"):
        a, b = np.load(tmp_path / "fused" / d / "sub4.npy"), np.load(chain / d / "sub4.npy")
        assert a.dtype == b.dtype and a.shape == b.shape, d
        assert np.array_equal(a, b), d


@pytest.mark.parametrize("length", (100, 200, 400))
def test_device_arithmetic_is_bit_identical_to_the_host_emulation(hostemu, length):
    """tests/hostemu compiles bandpower.cuh for the host with the scalar backend of cplx.cuh; the device runs the packed
    f32x2 backend.  Same operations in the same order with IEEE round-to-nearest -> the PSD values must agree bit for
    bit, which is what makes the CPU tier's error figures the GPU's."""
    rng = np.random.default_rng(length + 1)
    x = (30 * rng.standard_normal((4096, length)) + rng.uniform(-50, 50, (4096, 1))).astype(np.float32)
    _, psd_emu = hostemu(x)                                           # float64 view of float32 values
    de, psd = frontend.de_psd_windows(torch.from_numpy(x).to(DEV))
    assert np.array_equal(psd.cpu().numpy(), psd_emu.astype(np.float32))
