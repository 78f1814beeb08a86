"""CPU oracle for the EEG feature front end -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy / pure Python, the algorithm of the reference's
``EEG_preprocessing`` hot path (DE_PSD.py, segment_raw_signals_200Hz.py, segment_sliding_window.py and
the three ``extract_DE_PSD_features_*`` drivers).  It exists so that the CUDA path can be *checked*;
it is never the thing that is shipped or measured.

Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm.
Nothing under ``eeg2video_b200/`` imports it; the product path raises if the CUDA library is missing.

Parity status: the reference holds no tests, golden vectors or fixtures for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference``) and committed under ``tests/golden/``.
``tests/test_oracle_vs_reference.py`` additionally compares the oracle with the live reference whenever
``/root/reference`` is present.
"""
from .de_psd import (  # noqa: F401
    STFT_N, BAND_START_HZ, BAND_END_HZ, band_bin_ranges, hann_window,
    de_psd_loop, de_psd_closed_form,
    extract_de_psd_raw, extract_de_psd_1s, extract_de_psd_sw,
)
from .segment import (  # noqa: F401
    FS, clip_start, extract_2s_segment, segment_subject, seg_sliding_window, seq2seq_windows,
)
