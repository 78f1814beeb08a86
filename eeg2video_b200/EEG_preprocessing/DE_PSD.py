"""DE_PSD -- drop-in for /root/reference/EEG_preprocessing/DE_PSD.py:8-71 on the B200 kernel."""
import numpy as np

from .. import frontend
from . import _io


def DE_PSD(data, fre, time_window):
    '''
    compute DE and PSD (same contract as the reference function)
    --------
    input:  data [n*m]          n electrodes, m time points; m must equal int(time_window * fre)
            fre                 sampling rate (any positive rate; the drivers use 200)
            time_window         window length in seconds (any length >= 1 sample; the drivers use 0.5, 1 and 2)
    output: de, psd [n*5]       float64, five bands (delta, theta, alpha, beta, gamma) -- DE FIRST, like the
                                reference (DE_PSD.py:71)

    Hann window of int(time_window*fre) points, 200-point FFT (truncating / zero-padding, DE_PSD.py:58),
    band means of |X|^2 over bins range(int(f0 / fre * 200) - 1, int(f1 / fre * 200)) (DE_PSD.py:35-39, :63) --
    [0,3] [3,7] [7,13] [13,30] [30,98] at 200 Hz --, de = log2(100 * psd).  The three driver shapes (200 Hz; 100,
    200, 400 samples) run on the fused kernels, everything else on the general kernel (eegfe_de_psd_generic).
    Raises ValueError on a row-length mismatch (the reference's numpy broadcast error, DE_PSD.py:57),
    ValueError("math domain error") when a band has zero power (DE_PSD.py:68) and IndexError when a band reaches bin
    100 (fre < 198: the reference indexes past its 100 magnitudes, DE_PSD.py:64).
    '''
    if not fre > 0:
        raise ValueError(f"fre must be positive, got {fre!r}")
    like_torch = _io.is_torch(data)
    if data.ndim != 2:
        raise ValueError("data must be 2-D (electrodes, time points)")
    length = int(time_window * fre)
    if data.shape[1] != length:
        raise ValueError(
            f"operands could not be broadcast together with shapes ({data.shape[1]},) ({length},) ")
    if length < 1:
        raise ValueError(f"time_window * fre = {time_window * fre!r}: the window holds no sample")
    x = _io.to_device_f32(data)
    de, psd = frontend.de_psd_windows(x, check=True, fre=fre)
    return _io.finish((de, psd), like_torch, np.float64)
