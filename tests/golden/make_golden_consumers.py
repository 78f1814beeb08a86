"""Golden fixture for the consumer-side input build (next rows), made by RUNNING THE REFERENCE's own lines.

    python tests/golden/make_golden_consumers.py     # rewrites tests/golden/consumers_golden.npz

What runs here is the reference's code, not ours: ``GT_label`` and the ``Dataset`` class (whose constructor applies
``preprocessing.StandardScaler().fit(eeg)`` / ``transform``) are imported from the unmodified
``/root/reference/EEG2Video_New/Generation/models/train_semantic_predictor.py``; the ``__main__`` lines of that file
(:86-95, :114), of ``EEG2Video_New/Semantic/eeg_text.py`` (:115-125, :142-144) and of
``EEG-VP/EEG_VP_train_test.py`` (:232-267) cannot be imported (they sit under ``if __name__ == '__main__'`` or run at
import time on files that do not exist here), so they are quoted below with einops / sklearn exactly as written.
Inputs are int16 codes (value = 15 + code / 1024, exact in float32) on a reduced channel count (6 instead of 62).
"""
import importlib.util
import os

import numpy as np
import torch
from einops import rearrange
from sklearn import preprocessing
from sklearn.preprocessing import StandardScaler

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("EEG2VIDEO_REFERENCE", "/root/reference")


def decode(codes, dtype):
    return (15.0 + codes.astype(np.float64) / 1024.0).astype(dtype)


def main():
    spec = importlib.util.spec_from_file_location(
        "ref_tsp", os.path.join(REFERENCE_ROOT, "EEG2Video_New/Generation/models/train_semantic_predictor.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    GT_label = ref.GT_label
    chosed_label = [i for i in range(1, 41)]                                   # :77
    subset = [1, 10, 12, 16, 19, 23, 25, 31, 34, 39]                           # :76 (commented alternative)
    rng = np.random.default_rng(515)
    n_ch = 6
    codes_1s = rng.integers(-6000, 6000, (7, 40, 5, 2, n_ch, 5)).astype(np.int16)
    codes_1s[..., 3, 2] = 1234                                                 # one constant column (scale_ -> 1)
    codes_2s = rng.integers(-6000, 6000, (7, 40, 5, n_ch, 5)).astype(np.int16)
    out = {"gt_label": GT_label.astype(np.int32), "codes_1s": codes_1s, "codes_2s": codes_2s,
           "subset": np.array(subset, dtype=np.int32)}

    # ---- semantic predictor, 1 s features (float64 file): train_semantic_predictor.py:86-95, :114, Dataset :47-48 ----
    for tag, chosen in (("all", chosed_label), ("subset", subset)):
        eegdata = decode(codes_1s, np.float64)
        EEG = []
        for i in range(6):
            indices = [list(GT_label[i]).index(element) for element in chosen]
            chosed_eeg = eegdata[i][indices, :]
            EEG.append(chosed_eeg)
        EEG = np.stack(EEG, axis=0)
        EEG = torch.from_numpy(EEG)
        EEG = rearrange(EEG, 'a b c d e f -> (a b c) d (e f)')
        EEG = torch.mean(EEG, dim=1).resize(EEG.shape[0], n_ch * 5)
        dataset = ref.Dataset(EEG, np.zeros((EEG.shape[0], 1)))
        out[f"semantic_1s_{tag}"] = np.asarray(dataset.eeg)[::3].astype(np.float32)     # every 3rd row, float32

    # ---- semantic predictor, 2 s features (float32 file): eeg_text.py:115-125, :142-144 ----
    eegdata = decode(codes_2s, np.float32)
    eeg = []
    for i in range(6):
        indices = [list(GT_label[i]).index(element) for element in chosed_label]
        chosed_eeg = eegdata[i][indices, :]
        eeg.append(chosed_eeg)
    eeg = np.stack(eeg, axis=0)
    eeg = torch.from_numpy(eeg)
    eeg = rearrange(eeg, 'a b c e f -> (a b c) (e f)')
    normalize = preprocessing.StandardScaler()
    normalize.fit(eeg)
    out["semantic_2s_mean"] = normalize.mean_
    out["semantic_2s_var"] = normalize.var_
    out["semantic_2s_scale"] = normalize.scale_
    out["semantic_2s"] = np.asarray(normalize.transform(eeg))[::3]                # float32 (sklearn keeps the dtype)

    # ---- classifier folds: EEG_VP_train_test.py:232-267 (test_set_id = 0 and 3), every 16th row kept as float32 ----
    load_npy = decode(codes_1s, np.float64)
    All_train = rearrange(load_npy, 'a b c d e f -> a (b c d) e f')
    for test_set_id in (0, 3):
        val_set_id = test_set_id - 1
        if (val_set_id < 0):
            val_set_id = 6
        train_data = np.empty((0, n_ch, 5))
        for i in range(7):
            if (i == test_set_id):
                continue
            train_data = np.concatenate((train_data, All_train[i].reshape(400, n_ch, 5)))
        test_data = All_train[test_set_id]
        val_data = All_train[val_set_id]
        train_data = train_data.reshape(train_data.shape[0], n_ch * 5)
        test_data = test_data.reshape(test_data.shape[0], n_ch * 5)
        val_data = val_data.reshape(val_data.shape[0], n_ch * 5)
        normalize = StandardScaler()
        normalize.fit(train_data)
        train_data = normalize.transform(train_data)
        normalize = StandardScaler()
        normalize.fit(test_data)
        test_data = normalize.transform(test_data)
        normalize = StandardScaler()
        normalize.fit(val_data)
        val_data = normalize.transform(val_data)
        out[f"fold{test_set_id}_train"] = train_data[::16].astype(np.float32)
        out[f"fold{test_set_id}_test"] = test_data[::16].astype(np.float32)
        out[f"fold{test_set_id}_val"] = val_data[::16].astype(np.float32)

    path = os.path.join(HERE, "consumers_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
