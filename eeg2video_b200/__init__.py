"""eeg2video_b200 -- B200-native EEG feature front end (segmentation + Hann/FFT + five-band DE/PSD).

A from-scratch replacement for the hot path of gaspachoo/EEG2Video's ``EEG_preprocessing`` package:

* ``eeg2video_b200.EEG_preprocessing``  -- the reference's modules and call signatures (drop-in);
* ``eeg2video_b200.frontend``           -- fused, device-resident entry points (raw recordings in, features out);
* ``eeg2video_b200.pipeline``           -- host-resident recordings -> features (chunked H2D / kernel / D2H);
* ``eeg2video_b200.cohort``             -- subject sharding over the GPUs of one box + NCCL gather;
* ``eeg2video_b200.preprocess_all``     -- raw sub{N}.npy -> every feature directory of the reference in one pass;
* ``eeg2video_b200.glmnet_inputs``, ``consumers`` -- "next" rows: GLMNet input build, consumer-side scaler / re-ordering;
* ``eeg2video_b200.ops``                -- torch.library custom ops over the C ABI (include/eegfe.h);
* ``eeg2video_b200/csrc``               -- the CUDA kernels (sm_100a).

The CUDA library is required: there is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
