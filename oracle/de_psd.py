"""Oracle restatement of the DE/PSD band-power extraction (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/EEG_preprocessing/DE_PSD.py:8-71 and the drivers
extract_DE_PSD_features_1per2s.py:16-28, extract_DE_PSD_features_1per1s.py:24-58,
extract_DE_PSD_features_1per500ms.py:12-29.

Two formulations are given:

* ``de_psd_loop``        -- a scalar, loop-for-loop restatement (same float64 operation order as the
                            reference: Hann list, per-row FFT, per-band running sum, math.log(.., 2)).
                            It is the "port" that ``bench.py`` times as the CPU baseline.
* ``de_psd_closed_form`` -- the same mathematics vectorised over any number of rows (numpy rfft, float64);
                            used where the loop form would take minutes (full-size parity samples).
"""
import math

import numpy as np
from scipy.fftpack import fft as _fftpack_fft   # the reference's FFT (DE_PSD.py:5, call site :58)

STFT_N = 200                       # DE_PSD.py:27 -- FFT length is fixed, whatever the window length
BAND_START_HZ = (1, 4, 8, 14, 31)  # DE_PSD.py:28
BAND_END_HZ = (4, 8, 14, 31, 99)   # DE_PSD.py:29  (delta, theta, alpha, beta, gamma)


def band_bin_ranges(fre, stft_n=STFT_N):
    """Inclusive FFT-bin ranges ``[(lo, hi), ...]`` of the five bands.

    DE_PSD.py:35-39 computes ``int(f / fs * STFTN)`` in Python floats and DE_PSD.py:63 iterates
    ``range(fStartNum - 1, fEndNum)``; the divisor at :66 is ``fEndNum - fStartNum + 1``.
    For fre = 200 this is (0,3) (3,7) (7,13) (13,30) (30,98) with divisors 4, 5, 7, 18, 69.
    """
    out = []
    for f0, f1 in zip(BAND_START_HZ, BAND_END_HZ):
        s = int(f0 / fre * stft_n)
        e = int(f1 / fre * stft_n)
        out.append((s - 1, e - 1))
    return out


def hann_window(length):
    """DE_PSD.py:49-51: h[i] = 0.5 - 0.5*cos(2*pi*(i+1)/(L+1)), i = 0..L-1, float64."""
    k = np.arange(1, length + 1)
    return np.array([0.5 - 0.5 * np.cos(2 * np.pi * n / (length + 1)) for n in k])


def de_psd_loop(data, fre, time_window):
    """Loop-for-loop restatement of DE_PSD (DE_PSD.py:8-71).  Returns ``(de, psd)``, each (n, 5) float64."""
    data = np.asarray(data)
    rows = data.shape[0]
    ranges = band_bin_ranges(fre)
    window = hann_window(int(time_window * fre))
    psd = np.zeros((rows, len(ranges)))
    de = np.zeros((rows, len(ranges)))
    for row in range(rows):
        tapered = data[row] * window                       # :57 (raises ValueError on length mismatch)
        mag = abs(_fftpack_fft(tapered, STFT_N)[: STFT_N // 2])   # :58-59 (truncate or zero-pad to 200)
        for band, (lo, hi) in enumerate(ranges):
            energy = 0
            for k in range(lo, hi + 1):                    # :63-64
                energy = energy + mag[k] * mag[k]
            energy = energy / (hi - lo + 1)                # :66
            psd[row][band] = energy
            de[row][band] = math.log(100 * energy, 2)      # :68 (ValueError "math domain error" when 0)
    return de, psd


def de_psd_closed_form(windows, fre=200, time_window=None):
    """Vectorised float64 form of the same mathematics for an array ``(..., L)`` of windows.

    y = x * h_L ; X = FFT_200(y[..., :200] zero-padded) ; P = |X|^2 ; psd_b = mean(P[lo_b..hi_b]) ;
    de_b = log2(100 * psd_b).  Returns ``(de, psd)`` with shape ``(..., 5)``.
    Zero power yields -inf in ``de`` (the loop form raises instead, like the reference).
    """
    x = np.asarray(windows, dtype=np.float64)
    length = x.shape[-1]
    if time_window is not None and int(time_window * fre) != length:
        raise ValueError(f"operands could not be broadcast together with shapes ({length},) ({int(time_window * fre)},)")
    y = x * hann_window(length)
    spec = np.fft.rfft(y[..., :STFT_N], n=STFT_N, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2
    ranges = band_bin_ranges(fre)
    psd = np.stack([power[..., lo:hi + 1].sum(axis=-1) / (hi - lo + 1) for lo, hi in ranges], axis=-1)
    with np.errstate(divide="ignore"):
        de = np.log2(100.0 * psd)
    return de, psd


def _loop_or_closed(segment, fre, time_window, closed):
    if closed:
        return de_psd_closed_form(segment, fre, time_window)
    return de_psd_loop(segment, fre, time_window)


def extract_de_psd_raw(raw, fs=200, closed=True):
    """2 s driver (extract_DE_PSD_features_1per2s.py:16-28): (B,C,R,ch,400) -> two float32 (B,C,R,ch,5)."""
    raw = np.asarray(raw)
    shape = raw.shape[:4] + (5,)
    de_out = np.zeros(shape, dtype=np.float32)
    psd_out = np.zeros(shape, dtype=np.float32)
    for b in range(raw.shape[0]):
        for c in range(raw.shape[1]):
            for r in range(raw.shape[2]):
                de, psd = _loop_or_closed(raw[b, c, r], fs, 2, closed)
                de_out[b, c, r] = de
                psd_out[b, c, r] = psd
    return de_out, psd_out


def extract_de_psd_1s(raw, fs=200, closed=True):
    """1 s script body (extract_DE_PSD_features_1per1s.py:24-58) as a callable:
    (B,C,R,ch,400) -> two float64 (B,C,R,2,ch,5); window k covers samples [200k, 200k+200)."""
    raw = np.asarray(raw)
    shape = raw.shape[:3] + (2, raw.shape[3], 5)
    de_out = np.zeros(shape, dtype=np.float64)
    psd_out = np.zeros(shape, dtype=np.float64)
    for b in range(raw.shape[0]):
        for c in range(raw.shape[1]):
            for r in range(raw.shape[2]):
                for k in range(2):
                    de, psd = _loop_or_closed(raw[b, c, r, :, k * fs:(k + 1) * fs], fs, 1, closed)
                    de_out[b, c, r, k] = de
                    psd_out[b, c, r, k] = psd
    return de_out, psd_out


def extract_de_psd_sw(raw, fs, win_sec, closed=True):
    """500 ms driver (extract_DE_PSD_features_1per500ms.py:12-29):
    (B,C,R,W,ch,L) -> two float32 (B,C,R,W,ch,5)."""
    raw = np.asarray(raw)
    shape = raw.shape[:5] + (5,)
    de_out = np.zeros(shape, dtype=np.float32)
    psd_out = np.zeros(shape, dtype=np.float32)
    for b in range(raw.shape[0]):
        for c in range(raw.shape[1]):
            for r in range(raw.shape[2]):
                for w in range(raw.shape[3]):
                    de, psd = _loop_or_closed(raw[b, c, r, w], fs, win_sec, closed)
                    de_out[b, c, r, w] = de
                    psd_out[b, c, r, w] = psd
    return de_out, psd_out
