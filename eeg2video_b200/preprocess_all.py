"""One pass from raw recordings to every feature directory of the reference pipeline.

The reference runs five scripts in sequence, each reading and writing ``.npy`` files
(segment_raw_signals_200Hz -> segment_sliding_window -> extract_DE_PSD_features_1per2s / _1per1s / _1per500ms;
README.md "EEG preprocessing").  Here one command reads each ``eeg_root/sub{N}.npy`` once, uploads it once and writes

    DE_1per2s/  PSD_1per2s/  DE_1per1s/  PSD_1per1s/  DE_500ms_sw/  PSD_500ms_sw/        (always)
    Segmented_Rawf_200Hz_2s/  Segmented_500ms_sw/                                       (with --keep-segments)

under ``out_root`` with the reference's file names, shapes and dtypes (float32; float64 for the 1 s features), so
every downstream consumer reads the same files.  The intermediate clip / window tensors are never built unless asked for.

    python -m eeg2video_b200.preprocess_all --eeg_root ./data/EEG --out_root ./data/Preprocessing [--subs 1 2 3]

Several GPUs: subjects are independent, so there is nothing to exchange -- launch one process per GPU and every rank
takes every WORLD_SIZE-th recording on its own device (no process group is created):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m eeg2video_b200.preprocess_all ...
"""
import argparse
import os

import numpy as np
import torch

from . import frontend
from .EEG_preprocessing import _io, segment_raw_signals_200Hz, segment_sliding_window

FEATURE_DIRS = {"2s": ("DE_1per2s", "PSD_1per2s", np.float32),
                "1s": ("DE_1per1s", "PSD_1per1s", np.float64),
                "500ms": ("DE_500ms_sw", "PSD_500ms_sw", np.float32)}


def process_recording(raw, modes=("2s", "1s", "500ms")):
    """raw: numpy / torch recording (7, channels, T) -> {mode: (de, psd)} numpy arrays in the reference's layout and
    dtype.  The recording is uploaded once; every mode is one fused kernel launch over it."""
    dev = _io.to_device_f32(raw)
    out = {}
    for mode in modes:
        de, psd = frontend.de_psd_from_raw(dev, mode)
        dtype = FEATURE_DIRS[mode][2]
        out[mode] = (de.cpu().numpy().astype(dtype, copy=False), psd.cpu().numpy().astype(dtype, copy=False))
    return out, dev


def shard_for_this_rank(names, env=None):
    """The recordings this process handles when launched by torchrun (RANK / WORLD_SIZE in the environment):
    names[rank::world]; all of them when run alone."""
    env = os.environ if env is None else env
    world, rank = int(env.get("WORLD_SIZE", "1")), int(env.get("RANK", "0"))
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad RANK / WORLD_SIZE: {rank} / {world}")
    return list(names)[rank::world]


def preprocess_all(eeg_root="./data/EEG", out_root="./data/Preprocessing", subs=None, keep_segments=False, log=print):
    """Returns the list of file names written (by this rank)."""
    names = sorted(n for n in os.listdir(eeg_root) if n.endswith(".npy")) if subs is None \
        else [f"sub{int(s)}.npy" for s in subs]
    names = shard_for_this_rank(names)
    if torch.cuda.is_available() and "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())
    done = []
    for name in names:
        recording = np.load(os.path.join(eeg_root, name), mmap_mode="r")
        features, dev = process_recording(recording[:7])
        for mode, (de, psd) in features.items():
            for sub_dir, array in zip(FEATURE_DIRS[mode][:2], (de, psd)):
                os.makedirs(os.path.join(out_root, sub_dir), exist_ok=True)
                np.save(os.path.join(out_root, sub_dir, name), array)
        if keep_segments:
            # the intermediates keep the recording's own element type, like the reference's files (byte-exact gathers)
            clips = segment_raw_signals_200Hz.segment_subject(np.asarray(recording[:7]))
            windows = segment_sliding_window.materialize_windows(clips)
            for sub_dir, array in (("Segmented_Rawf_200Hz_2s", clips), ("Segmented_500ms_sw", windows)):
                os.makedirs(os.path.join(out_root, sub_dir), exist_ok=True)
                np.save(os.path.join(out_root, sub_dir, name), array)
            del clips, windows
        del dev
        torch.cuda.empty_cache()
        log(f"{name}: " + ", ".join(f"{m} {features[m][0].shape}" for m in features))
        done.append(name)
    return done


def main(argv=None):
    cli = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    cli.add_argument("--eeg_root", default="./data/EEG")
    cli.add_argument("--out_root", default="./data/Preprocessing")
    cli.add_argument("--subs", nargs="+", type=int, default=None, help="subject numbers (default: every sub*.npy found)")
    cli.add_argument("--keep-segments", action="store_true", help="also write the Segmented_* intermediates")
    opt = cli.parse_args(argv)
    return preprocess_all(opt.eeg_root, opt.out_root, opt.subs, opt.keep_segments)


if __name__ == "__main__":
    main()
