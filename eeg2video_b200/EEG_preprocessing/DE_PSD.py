"""DE_PSD -- drop-in for /root/reference/EEG_preprocessing/DE_PSD.py:8-71 on the B200 kernel."""
import numpy as np

from .. import frontend
from . import _io

_WINDOW_SECONDS = {100: 0.5, 200: 1, 400: 2}


def DE_PSD(data, fre, time_window):
    '''
    compute DE and PSD (same contract as the reference function)
    --------
    input:  data [n*m]          n electrodes, m time points; m must equal int(time_window * fre)
            fre                 sampling rate (200)
            time_window         window length in seconds (0.5, 1 or 2)
    output: de, psd [n*5]       float64, five bands (delta, theta, alpha, beta, gamma) -- DE FIRST, like the
                                reference (DE_PSD.py:71)

    Hann window of int(time_window*fre) points, 200-point FFT (truncating / zero-padding, DE_PSD.py:58),
    band means of |X|^2 over bins [0,3] [3,7] [7,13] [13,30] [30,98], de = log2(100 * psd).
    Raises ValueError on a row-length mismatch (the reference's numpy broadcast error, DE_PSD.py:57) and
    ValueError("math domain error") when a band has zero power (DE_PSD.py:68).
    '''
    _io.check_fs(fre, "fre")
    like_torch = _io.is_torch(data)
    if data.ndim != 2:
        raise ValueError("data must be 2-D (electrodes, time points)")
    length = int(time_window * fre)
    if data.shape[1] != length:
        raise ValueError(
            f"operands could not be broadcast together with shapes ({data.shape[1]},) ({length},) ")
    if length not in _WINDOW_SECONDS:
        raise NotImplementedError(f"time_window={time_window!r}: supported window lengths are 0.5, 1 and 2 s")
    x = _io.to_device_f32(data)
    de, psd = frontend.de_psd_windows(x, check=True)
    return _io.finish((de, psd), like_torch, np.float64)
