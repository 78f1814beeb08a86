"""Synthetic SEED-DV-shaped recordings (there is no dataset access): per subject float32 (7, 62, 104000) --
62 channels, 200 Hz, 7 blocks of 8 min 40 s.

    x = 30 * (0.7 * pink + 0.3 * white) + 10 * sin(2 pi 10 t / 200 + phi_ch) + offset_ch
    offset_ch ~ U(-50, 50), phi_ch ~ U(0, 2 pi), pink normalised to unit std per row, seed = 1000 + subject

(SURVEY.md section 8d / BASELINE.md section 3: microvolt-like scale, moderate DC, an alpha-band tone.)
Generated with torch on whichever device is asked for, so the same code feeds the GPU bench and -- through
``.cpu().numpy()`` -- the oracle.  torch.fft is used here for the 1/f shaping only; it is data generation, not
part of the measured path.
"""
import math

import torch

BLOCKS, CHANNELS, BLOCK_LEN, FS = 7, 62, 104000, 200
BYTES_PER_SUBJECT = BLOCKS * CHANNELS * BLOCK_LEN * 4


def synth_blocks(n_blocks, seed, device="cpu", channels=CHANNELS, block_len=BLOCK_LEN, kind="pink_tone"):
    """(n_blocks, channels, block_len) float32 on `device`; deterministic in (seed, kind, device type)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    shape = (n_blocks, channels, block_len)
    white = torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)
    if kind == "white":
        return white.mul_(30.0)
    if kind != "pink_tone":
        raise ValueError(kind)
    src = torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)
    spec = torch.fft.rfft(src, dim=-1)
    del src
    f = torch.arange(spec.shape[-1], device=dev, dtype=torch.float32)
    f[0] = 1.0
    spec.mul_(torch.rsqrt(f))
    pink = torch.fft.irfft(spec, n=block_len, dim=-1)
    del spec
    pink.div_(pink.std(dim=-1, keepdim=True))
    phase = torch.rand((n_blocks, channels, 1), generator=gen, device=dev) * (2 * math.pi)
    offset = torch.rand((n_blocks, channels, 1), generator=gen, device=dev) * 100.0 - 50.0
    t = torch.arange(block_len, device=dev, dtype=torch.float32)
    x = pink.mul_(0.7 * 30.0).add_(white.mul_(0.3 * 30.0))
    del white
    x.add_(10.0 * torch.sin((2 * math.pi * 10.0 / FS) * t + phase)).add_(offset)
    return x


def synth_subject(subject, device="cpu", kind="pink_tone"):
    """One subject: (7, 62, 104000) float32, seed 1000 + subject."""
    return synth_blocks(BLOCKS, 1000 + int(subject), device=device, kind=kind)


def synth_cohort(subjects, device, kind="pink_tone", out=None):
    """Stack of subjects: (len(subjects), 7, 62, 104000) float32 on `device` (generated one subject at a time)."""
    subjects = list(subjects)
    if out is None:
        out = torch.empty((len(subjects), BLOCKS, CHANNELS, BLOCK_LEN), dtype=torch.float32, device=device)
    for i, s in enumerate(subjects):
        out[i].copy_(synth_subject(s, device=device, kind=kind))
    return out
