"""Oracle for the consumer-side input build (TEST INFRASTRUCTURE; see oracle/__init__.py).

numpy restatement of what the reference's consumers do to the feature files before training
(EEG2Video_New/Generation/models/train_semantic_predictor.py:47-48, :87-95, :114; EEG2Video_New/Semantic/
eeg_text.py:115-125, :142-144; EEG-VP/EEG_VP_train_test.py:232-267) and of scikit-learn's StandardScaler, which does
the arithmetic there (third-party; not pinned by the reference's requirements.txt -- restated from scikit-learn 1.9.0:
preprocessing/_data.py StandardScaler.partial_fit / transform / _is_constant_feature, utils/extmath.py
_incremental_mean_and_var).  Pinned by tests/golden/consumers_golden.npz, which make_golden.py produces by running
the reference's own lines with sklearn itself.
"""
import numpy as np


def concept_order(gt_label_row, chosen):
    row = list(gt_label_row)
    return [row.index(e) for e in chosen]                       # train_semantic_predictor.py:88


def select_concepts(eegdata, gt_label, chosen, blocks=range(6)):
    """eegdata (blocks, 40, 5, ...) -> stacked (len(blocks), len(chosen), 5, ...)   (:86-91)"""
    out = []
    for i in blocks:
        indices = concept_order(gt_label[i], chosen)
        out.append(eegdata[i][indices, :])
    return np.stack(out, axis=0)


def standard_scaler_fit(x):
    """(mean_, var_, scale_) float64 as StandardScaler().fit(x) computes them."""
    x = np.asarray(x)
    n = x.shape[0]
    xs = x.astype(np.float64)
    new_sum = xs.sum(axis=0)
    mean = new_sum / n
    temp = xs - new_sum / n
    correction = temp.sum(axis=0)
    var = ((temp ** 2).sum(axis=0) - correction ** 2 / n) / n
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2      # _is_constant_feature
    scale = np.sqrt(var)
    scale[constant] = 1.0
    return mean, var, scale


def standard_scaler_transform(x, mean, scale):
    """float64 result, as at every call site of the reference: they pass float64 arrays or torch tensors, which
    scikit-learn converts to float64 before `X -= mean_; X /= scale_` (StandardScaler.transform)."""
    x = np.array(x, dtype=np.float64, copy=True)
    x -= mean
    x /= scale
    return x


def semantic_predictor_inputs(eegdata, gt_label, chosen, blocks=range(6)):
    """(N, 310) standardised EEG matrix of the semantic predictor (1 s features: mean over the two windows)."""
    eeg = select_concepts(eegdata, gt_label, chosen, blocks)
    if eeg.ndim == 6:                                            # a b c d e f -> (a b c) d (e f), mean over d  (:95, :114)
        a, b, c, d, e, f = eeg.shape
        eeg = eeg.reshape(a * b * c, d, e * f).mean(axis=1)
    else:                                                        # a b c e f -> (a b c) (e f)  (eeg_text.py:125)
        a, b, c, e, f = eeg.shape
        eeg = eeg.reshape(a * b * c, e * f)
    mean, _, scale = standard_scaler_fit(eeg)
    return standard_scaler_transform(eeg, mean, scale)


def classifier_fold_inputs(load_npy, test_block):
    """EEG_VP_train_test.py:232-267 for one fold."""
    a, b, c, d, e, f = load_npy.shape
    all_train = load_npy.reshape(a, b * c * d, e, f)
    val_block = test_block - 1 if test_block > 0 else a - 1
    train = np.concatenate([all_train[i].reshape(b * c * d, e, f) for i in range(a) if i != test_block])
    out = {}
    for name, data in (("train", train), ("test", all_train[test_block]), ("val", all_train[val_block])):
        data = data.reshape(data.shape[0], e * f)
        mean, _, scale = standard_scaler_fit(data)
        out[name] = standard_scaler_transform(data, mean, scale)
    return out
