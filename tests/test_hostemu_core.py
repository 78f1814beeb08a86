"""The CUDA kernel's arithmetic core (bandpower.cuh), compiled for the host by tests/hostemu, against the oracle.

This checks -- without a GPU -- the prime-factor index maps, radix-8 / radix-5 butterflies, the real-pair
separation, the band tables and every constant of the device code; the fp32 operation order is the device's,
so the error figures seen here are the ones the GPU produces.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from conftest import assert_features_close, decode

TW = {100: 0.5, 200: 1, 400: 2}


@pytest.mark.parametrize("length", (100, 200, 400))
@pytest.mark.parametrize("kind", ("white", "offset", "pink_tone", "small"))
def test_core_matches_reference_golden(hostemu, golden, length, kind):
    g = golden("de_psd_golden.npz")
    key = f"L{length}_{kind}"
    de, psd = hostemu(decode(g[key + "_codes"]))
    assert_features_close(de, psd, g[key + "_de"], g[key + "_psd"])


@pytest.mark.parametrize("length", (100, 200, 400))
def test_core_bulk_random(hostemu, length):
    rng = np.random.default_rng(length)
    x = (30 * rng.standard_normal((20000, length)) + rng.uniform(-50, 50, (20000, 1))).astype(np.float32)
    de_ref, psd_ref = oracle.de_psd_closed_form(x, 200, TW[length])
    de, psd = hostemu(x)
    assert_features_close(de, psd, de_ref, psd_ref)
    # typical error is two orders inside the bar
    assert np.max(np.abs(psd - psd_ref) / psd_ref) < 5e-6


def test_core_impulse_and_single_bins(hostemu):
    """Known answers: an impulse gives flat power; a pure on-grid tone concentrates in its band."""
    x = np.zeros((3, 100), np.float32)
    x[:, 0] = (1.0, 8.0, 1024.0)
    de, psd = hostemu(x)
    h0 = oracle.hann_window(100)[0]
    assert np.allclose(psd, ((x[:, 0].astype(np.float64) * h0) ** 2)[:, None], rtol=1e-5)
    t = np.arange(200)
    for f_hz, band in ((2, 0), (6, 1), (10, 2), (20, 3), (60, 4)):
        tone = np.cos(2 * np.pi * f_hz * t / 200).astype(np.float32)[None]
        de_ref, psd_ref = oracle.de_psd_closed_form(tone, 200, 1)
        de, psd = hostemu(tone)
        assert np.argmax(psd[0]) == band
        assert abs(psd[0, band] - psd_ref[0, band]) / psd_ref[0, band] < 1e-5


@pytest.mark.parametrize("length", (100, 200, 400))
def test_core_every_sample_position_matters_correctly(hostemu, length):
    """Impulse at every sample position: catches any wrong sample <-> (n1, n2) assignment or Hann index."""
    n = length
    x = np.zeros((n, length), np.float32)
    x[np.arange(n), np.arange(n)] = 100.0
    de, psd = hostemu(x)
    _, psd_ref = oracle.de_psd_closed_form(x, 200, TW[length])
    live = psd_ref.min(axis=1) > 0          # 2 s mode: samples 200..399 do not reach the FFT
    if length == 400:
        assert not live[200:].any() and np.all(psd[200:] == 0)
    if length != 400:
        # h is tiny at the window edges, keep the comparison relative
        assert np.max(np.abs(psd[live] - psd_ref[live]) / psd_ref[live]) < 1e-4
    else:
        sel = np.arange(200)
        assert np.max(np.abs(psd[sel] - psd_ref[sel]) / psd_ref[sel]) < 1e-4


@settings(max_examples=40, deadline=None)
@given(scale=st.floats(min_value=1e-3, max_value=1e4), offset=st.floats(min_value=-500, max_value=500),
       length=st.sampled_from((100, 200, 400)), seed=st.integers(0, 2 ** 31 - 1))
def test_core_amplitude_and_offset_sweep(hostemu, scale, offset, length, seed):
    rng = np.random.default_rng(seed)
    x = (scale * rng.standard_normal((64, length)) + offset * scale / 30.0).astype(np.float32)
    de_ref, psd_ref = oracle.de_psd_closed_form(x, 200, TW[length])
    de, psd = hostemu(x)
    assert_features_close(de, psd, de_ref, psd_ref)


def test_core_power_of_two_scaling_is_exact(hostemu):
    """|X|^2 is quadratic: scaling the input by 2 scales every band energy by exactly 4 in binary fp32."""
    rng = np.random.default_rng(5)
    x = (30 * rng.standard_normal((256, 100))).astype(np.float32)
    _, p1 = hostemu(x)
    _, p2 = hostemu(2 * x)
    assert np.array_equal(4 * p1, p2)


@pytest.mark.parametrize("length", (100, 200, 400))
def test_shifted_reads_feed_identical_arithmetic(hostemu, length):
    """The kernels' shifted-span form (rows that are not 16-byte aligned) reads the window 0..3 floats into its
    shared-memory row with scalar loads, or with 64-bit loads when the offset is even: same samples, same operations,
    hence the same bits as the aligned 64- / 128-bit read path -- and nothing outside the window is touched (the
    emulation poisons it)."""
    rng = np.random.default_rng(length)
    x = (30 * rng.standard_normal((64, length))).astype(np.float32)
    want = hostemu.band_energy(x)
    for shift in range(4):
        assert np.array_equal(hostemu.band_energy(x, shift, 1), want), (shift, 1)
        if shift % 2 == 0:
            assert np.array_equal(hostemu.band_energy(x, shift, 2), want), (shift, 2)
