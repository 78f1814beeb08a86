"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE (gaspachoo/EEG2Video).

The reference has no tests or golden vectors of its own for the EEG front end, so the only
way to pin the oracle and the CUDA path is against outputs of the reference code itself.  This script imports
the unmodified modules from ``/root/reference/EEG_preprocessing`` (build container only -- the GPU box does not
have them), feeds them seeded inputs and stores inputs + outputs as small ``.npz`` files.

Inputs are stored as int16 codes; the float32 signal is ``codes * scale`` with a power-of-two scale, so it is
reproduced exactly on any machine.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import contextlib
import io
import os
import runpy
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("EEG2VIDEO_REFERENCE", "/root/reference")
SCALE = np.float32(2.0 ** -5)      # int16 code -> microvolt-like float32, exact


def signal_codes(rng, shape, kind):
    """Seeded int16 codes for a few EEG-like signal families (last axis = time, fs = 200)."""
    n = shape[-1]
    t = np.arange(n)
    white = rng.standard_normal(shape)
    if kind == "white":                     # sigma 30
        x = 30.0 * white
    elif kind == "offset":                  # sigma 30 with per-row DC offset in [-50, 50]
        x = 30.0 * white + rng.uniform(-50, 50, shape[:-1] + (1,))
    elif kind == "pink_tone":               # bench distribution: 1/f-ish + white + 10 Hz tone + offset
        spec = np.fft.rfft(rng.standard_normal(shape), axis=-1)
        f = np.arange(spec.shape[-1], dtype=np.float64)
        f[0] = 1.0
        pink = np.fft.irfft(spec / np.sqrt(f), n=n, axis=-1)
        pink /= pink.std(axis=-1, keepdims=True)
        phase = rng.uniform(0, 2 * np.pi, shape[:-1] + (1,))
        x = 30.0 * (0.7 * pink + 0.3 * white) + 10.0 * np.sin(2 * np.pi * 10 * t / 200 + phase) \
            + rng.uniform(-50, 50, shape[:-1] + (1,))
    elif kind == "small":                   # tiny amplitudes -> negative DE
        x = 0.2 * white
    else:
        raise ValueError(kind)
    return np.clip(np.round(x / float(SCALE)), -32768, 32767).astype(np.int16)


def decode(codes):
    return codes.astype(np.float32) * SCALE


def main():
    sys.path.insert(0, REFERENCE_ROOT)
    from EEG_preprocessing.DE_PSD import DE_PSD
    from EEG_preprocessing.segment_raw_signals_200Hz import extract_2s_segment
    from EEG_preprocessing.segment_sliding_window import seg_sliding_window
    from EEG_preprocessing.extract_DE_PSD_features_1per2s import extract_de_psd_raw
    from EEG_preprocessing.extract_DE_PSD_features_1per500ms import extract_de_psd_sw

    # ---- 1. DE_PSD known answers: three window lengths x four signal families, 62 rows each -------------
    rng = np.random.default_rng(20240607)
    out = {}
    for length, tw in ((400, 2), (200, 1), (100, 0.5)):
        for kind in ("white", "offset", "pink_tone", "small"):
            codes = signal_codes(rng, (62, length), kind)
            de, psd = DE_PSD(decode(codes), 200, tw)
            key = f"L{length}_{kind}"
            out[key + "_codes"] = codes
            out[key + "_de"] = de
            out[key + "_psd"] = psd
    # analytic case: unit impulse at sample 0 -> every bin has power h[0]^2
    imp = np.zeros((3, 100), dtype=np.float32)
    imp[:, 0] = (1.0, 8.0, 1024.0)
    de, psd = DE_PSD(imp, 200, 0.5)
    out["impulse_x"] = imp
    out["impulse_de"] = de
    out["impulse_psd"] = psd
    np.savez_compressed(os.path.join(HERE, "de_psd_golden.npz"), **out)

    # ---- 2. segmentation: formula-defined raw block, selected (block, concept, repetition) ---------------
    # raw[b, ch, t] = ((b * 7919 + ch * 104729 + t * 31) mod 65536) - 32768 as int16 (regenerated in the test)
    n_ch, n_t = 5, 104000
    b = np.arange(7).reshape(7, 1, 1)
    ch = np.arange(n_ch).reshape(1, n_ch, 1)
    t = np.arange(n_t).reshape(1, 1, n_t)
    raw = (((b * 7919 + ch * 104729 + t * 31) % 65536) - 32768).astype(np.int16)
    picks = [(0, 0, 0), (0, 0, 4), (3, 17, 2), (6, 39, 4), (6, 39, 0), (1, 1, 1), (5, 20, 3)]
    segs = np.stack([np.array(extract_2s_segment(block=bb, concept=cc, repetition=rr, data=raw))
                     for bb, cc, rr in picks])
    np.savez_compressed(os.path.join(HERE, "segment_golden.npz"), picks=np.array(picks), segments=segs,
                        n_ch=n_ch, n_t=n_t)

    # ---- 3. drivers on a small clip tensor (2 blocks x 2 concepts x 3 reps x 62 ch x 400) -----------------
    rng = np.random.default_rng(7)
    codes = signal_codes(rng, (2, 2, 3, 62, 400), "pink_tone")
    clips = decode(codes)
    de2, psd2 = extract_de_psd_raw(clips, 200)
    win = seg_sliding_window(clips, 0.5, 0.25, fs=200)
    de5, psd5 = extract_de_psd_sw(np.ascontiguousarray(win), 200, 0.5)
    # the 1 s path is a module-level script in the reference; its inner call pattern (:46-47) on each half:
    de1 = np.zeros((2, 2, 3, 2, 62, 5))
    psd1 = np.zeros_like(de1)
    for i in range(2):
        for j in range(2):
            for k in range(3):
                for h in range(2):
                    d, p = DE_PSD(clips[i, j, k, :, h * 200:(h + 1) * 200].reshape(62, 200), 200, 1)
                    de1[i, j, k, h], psd1[i, j, k, h] = d, p
    np.savez_compressed(os.path.join(HERE, "drivers_golden.npz"), codes=codes,
                        de_2s=de2, psd_2s=psd2, de_1s=de1, psd_1s=psd1, de_500ms=de5, psd_500ms=psd5,
                        window_shape=np.array(win.shape), window_strides=np.array(win.strides))

    # ---- 4. the real 1 s script, run with runpy on one full-size seeded subject; keep a sample of outputs --
    rng = np.random.default_rng(11)
    full = signal_codes(rng, (7, 40, 5, 62, 400), "offset")
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data/Preprocessing/Segmented_Rawf_200Hz_2s"))
        np.save(os.path.join(tmp, "data/Preprocessing/Segmented_Rawf_200Hz_2s/sub1.npy"), decode(full))
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                runpy.run_path(os.path.join(REFERENCE_ROOT, "EEG_preprocessing/extract_DE_PSD_features_1per1s.py"))
        finally:
            os.chdir(cwd)
        de = np.load(os.path.join(tmp, "data/Preprocessing/DE_1per1s/sub1.npy"))
        psd = np.load(os.path.join(tmp, "data/Preprocessing/PSD_1per1s/sub1.npy"))
    pick_rng = np.random.default_rng(3)
    idx = np.stack([pick_rng.integers(0, 7, 12), pick_rng.integers(0, 40, 12), pick_rng.integers(0, 5, 12)], axis=1)
    np.savez_compressed(os.path.join(HERE, "script_1s_golden.npz"), seed=11, kind="offset", idx=idx,
                        out_shape=np.array(de.shape), out_dtype=str(de.dtype),
                        clips_codes=np.stack([full[i, j, k] for i, j, k in idx]),
                        de=np.stack([de[i, j, k] for i, j, k in idx]),
                        psd=np.stack([psd[i, j, k] for i, j, k in idx]))
    for name in sorted(os.listdir(HERE)):
        if name.endswith(".npz"):
            print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


if __name__ == "__main__":
    main()
