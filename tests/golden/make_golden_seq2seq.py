"""Golden vector for the Seq2Seq window layout, made by EXECUTING the reference's own lines
(EEG2Video_New/Seq2Seq/my_autoregressive_transformer.py:309-314 -- they sit inside `if __name__ == "__main__"`, so they
are cut out of the file by line number and run on a small seeded tensor).  Build container only:

    python tests/golden/make_golden_seq2seq.py        # writes tests/golden/seq2seq_golden.npz
"""
import os
import textwrap

import numpy as np
import torch

REF = os.environ.get("EEG2VIDEO_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "EEG2Video_New", "Seq2Seq", "my_autoregressive_transformer.py")


def reference_windows(new_eeg):
    """Run lines 309-314 of the reference file on `new_eeg` (a torch tensor (..., ch, 400)); returns EEG."""
    with open(SRC) as f:
        lines = f.readlines()[308:314]
    code = textwrap.dedent("".join(lines))
    assert "window_size = 100" in code and "torch.stack(EEG, dim=-1)" in code, "reference lines moved"
    scope = {"torch": torch, "new_eeg": new_eeg}
    exec(compile(code, SRC, "exec"), scope)
    return scope["EEG"]


if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    codes = rng.integers(-3000, 3000, (2, 3, 5, 400)).astype(np.int16)
    out = reference_windows(torch.from_numpy(codes.astype(np.float32))).numpy()
    assert out.shape == (2, 3, 5, 100, 7)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "seq2seq_golden.npz"),
                        codes=codes, windows=out.astype(np.int16))
    print("wrote seq2seq_golden.npz", out.shape)
