"""Consumer-side input build: concept re-ordering, window mean, 310-column flatten,
StandardScaler.  CPU tier: the oracle against the golden fixture made by the reference's own lines
(tests/golden/make_golden_consumers.py) and against scikit-learn itself.  GPU tier: the CUDA kernels (through the
C ABI) against the same fixture and the oracle.

Tolerances (floating point): the reference runs the scaler in float64 at every call site; the CUDA path keeps the
statistics in float64 and rounds the result once to float32 -> bar 1 ulp(float32) of the reference value when the
inputs are the same float32 numbers, 5e-6 absolute on standardised (O(1)) values for the window-mean path (float32
mean of two float32 features against the reference's float64 mean).
"""
import numpy as np
import pytest

from oracle import consumers as oc

F32_OUT_ATOL = 5e-6


def decode(codes, dtype):
    return (15.0 + codes.astype(np.float64) / 1024.0).astype(dtype)


@pytest.fixture(scope="module")
def gold(golden):
    return golden("consumers_golden.npz")


# ---- CPU: oracle pinned to the reference's outputs ------------------------------------------------------------------
def test_oracle_semantic_inputs_match_reference_lines(gold):
    gt = gold["gt_label"]
    eeg1 = decode(gold["codes_1s"], np.float64)
    for tag, chosen in (("all", list(range(1, 41))), ("subset", [int(v) for v in gold["subset"]])):
        got = oc.semantic_predictor_inputs(eeg1, gt, chosen)
        assert got.dtype == np.float64
        assert np.allclose(got[::3], gold[f"semantic_1s_{tag}"], rtol=0, atol=1e-6)        # fixture stored as float32
    eeg2 = decode(gold["codes_2s"], np.float32)
    sel = oc.select_concepts(eeg2, gt, list(range(1, 41)))
    flat = sel.reshape(-1, sel.shape[-2] * sel.shape[-1])
    mean, var, scale = oc.standard_scaler_fit(flat)
    assert np.allclose(mean, gold["semantic_2s_mean"], rtol=1e-13, atol=0)
    assert np.allclose(var, gold["semantic_2s_var"], rtol=1e-11, atol=0)
    assert np.allclose(scale, gold["semantic_2s_scale"], rtol=1e-11, atol=0)
    got = oc.semantic_predictor_inputs(eeg2, gt, list(range(1, 41)))
    assert got.dtype == np.float64 and np.allclose(got[::3], gold["semantic_2s"], rtol=1e-12, atol=1e-13)


def test_oracle_classifier_folds_match_reference_lines(gold):
    eeg1 = decode(gold["codes_1s"], np.float64)
    for fold in (0, 3):
        got = oc.classifier_fold_inputs(eeg1, fold)
        for name in ("train", "test", "val"):
            assert np.allclose(got[name][::16], gold[f"fold{fold}_{name}"], rtol=0, atol=1e-6)


def test_oracle_scaler_equals_sklearn():
    sklearn = pytest.importorskip("sklearn.preprocessing")
    rng = np.random.default_rng(2)
    for dtype in (np.float32, np.float64):
        x = (20 + 3 * rng.standard_normal((777, 31))).astype(dtype)
        x[:, 4] = 7.25                                           # constant column
        x[:, 9] *= 1e-3
        ref = sklearn.StandardScaler().fit(x)
        mean, var, scale = oc.standard_scaler_fit(x)
        assert np.allclose(mean, ref.mean_, rtol=1e-13, atol=0) and np.allclose(var, ref.var_, rtol=1e-10, atol=1e-30)
        assert np.array_equal(scale == 1.0, ref.scale_ == 1.0) and np.allclose(scale, ref.scale_, rtol=1e-10)
        import torch
        got = oc.standard_scaler_transform(x, ref.mean_, ref.scale_)
        want = ref.transform(torch.from_numpy(x))                # the reference passes torch tensors (or float64 arrays)
        assert got.dtype == np.float64 and want.dtype == np.float64 and np.array_equal(got, want)


def test_concept_order_is_list_index(gold):
    gt = gold["gt_label"]
    from eeg2video_b200 import consumers
    chosen = [int(v) for v in gold["subset"]]
    for b in range(7):
        assert consumers.concept_order(gt[b], chosen) == [list(gt[b]).index(e) for e in chosen]
    with pytest.raises(ValueError):
        consumers.concept_order(gt[0], [41])
    idx = consumers.clip_index([2, 0], gt, chosen)
    assert idx.dtype == np.int32 and idx.shape == (2 * len(chosen) * 5,)
    assert idx[0] == (2 * 40 + list(gt[2]).index(chosen[0])) * 5 and idx[4] == idx[0] + 4


def test_scaler_is_cuda_only():
    import torch
    from eeg2video_b200 import consumers
    with pytest.raises(RuntimeError, match="CUDA-only"):
        consumers.StandardScaler().fit(torch.zeros(4, 3))


# ---- GPU: the CUDA path -----------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_semantic_inputs_against_reference_fixture(gold):
    import torch
    from eeg2video_b200 import consumers
    gt = gold["gt_label"]
    feat1 = torch.from_numpy(decode(gold["codes_1s"], np.float32)).cuda()
    for tag, chosen in (("all", list(range(1, 41))), ("subset", [int(v) for v in gold["subset"]])):
        x, scaler = consumers.semantic_predictor_inputs(feat1, gt, chosen)
        assert x.dtype == torch.float32 and tuple(x.shape) == (6 * len(chosen) * 5, 30)
        assert np.max(np.abs(x.cpu().numpy()[::3] - gold[f"semantic_1s_{tag}"])) <= F32_OUT_ATOL
        assert float(scaler.scale_[3 * 5 + 2]) == 1.0            # the constant column keeps scale 1 (sklearn rule)
    feat2 = torch.from_numpy(decode(gold["codes_2s"], np.float32)).cuda()
    x, scaler = consumers.semantic_predictor_inputs(feat2, gt, list(range(1, 41)))
    assert np.allclose(scaler.mean_.cpu().numpy(), gold["semantic_2s_mean"], rtol=1e-13, atol=0)
    assert np.allclose(scaler.var_.cpu().numpy(), gold["semantic_2s_var"], rtol=1e-11, atol=0)
    want = gold["semantic_2s"]                                   # float64: sklearn converts the torch tensor
    got = x.cpu().numpy()[::3].astype(np.float64)
    assert np.all(np.abs(got - want) <= np.spacing(np.abs(want).astype(np.float32)))        # one rounding to float32


@pytest.mark.gpu
def test_gpu_classifier_folds_against_reference_fixture(gold):
    import torch
    from eeg2video_b200 import consumers
    feat = torch.from_numpy(decode(gold["codes_1s"], np.float32)).cuda()
    for fold in (0, 3):
        got = consumers.classifier_fold_inputs(feat, fold)
        assert tuple(got["train"].shape) == (2400, 30) and tuple(got["test"].shape) == (400, 30)
        for name in ("train", "test", "val"):
            assert np.max(np.abs(got[name].cpu().numpy()[::16] - gold[f"fold{fold}_{name}"])) <= F32_OUT_ATOL


@pytest.mark.gpu
@pytest.mark.parametrize("n_rows,n_cols", ((1, 1), (5, 310), (64, 310), (65, 3), (2400, 310), (4097, 17)))
def test_gpu_scaler_against_oracle(n_rows, n_cols):
    import torch
    from eeg2video_b200 import consumers
    rng = np.random.default_rng(n_rows * 31 + n_cols)
    x = (20 + 3 * rng.standard_normal((n_rows, n_cols))).astype(np.float32)
    if n_cols > 2:
        x[:, 1] = -3.5
    mean, var, scale = oc.standard_scaler_fit(x)
    sc = consumers.StandardScaler().fit(torch.from_numpy(x).cuda())
    assert np.allclose(sc.mean_.cpu().numpy(), mean, rtol=1e-13, atol=0)
    assert np.allclose(sc.var_.cpu().numpy(), var, rtol=1e-9, atol=1e-25)
    assert np.array_equal(sc.scale_.cpu().numpy() == 1.0, scale == 1.0)
    want = oc.standard_scaler_transform(x, mean, scale)
    got = sc.transform(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.all(np.abs(got.astype(np.float64) - want) <= np.spacing(np.abs(want).astype(np.float32)) + 1e-30)
    wide = torch.from_numpy(np.concatenate([x, x], axis=1)).cuda()                      # strided rows (a column slice)
    again = sc.transform(wide[:, :n_cols]).cpu().numpy()
    assert np.array_equal(again, got)


@pytest.mark.gpu
def test_gpu_select_clips_layouts():
    import torch
    from eeg2video_b200 import consumers
    rng = np.random.default_rng(8)
    gt = np.stack([rng.permutation(40) + 1 for _ in range(3)])
    feat = torch.from_numpy(rng.standard_normal((3, 40, 5, 7, 4, 5)).astype(np.float32)).cuda()
    chosen = [7, 1, 33]
    kept = consumers.select_clips(feat, [2, 0], gt, chosen)                             # windows kept
    assert tuple(kept.shape) == (2 * 3 * 5, 7, 20)
    ref = oc.select_concepts(feat.cpu().numpy()[[2, 0]], gt[[2, 0]], chosen, blocks=range(2))
    assert np.array_equal(kept.cpu().numpy(), ref.reshape(30, 7, 20))
    mean = consumers.select_clips(feat, [2, 0], gt, chosen, mean_windows=True)
    assert np.allclose(mean.cpu().numpy(), ref.reshape(30, 7, 20).mean(axis=1), rtol=0, atol=1e-6)
    empty = consumers.select_clips(feat, [], gt, chosen)
    assert tuple(empty.shape) == (0, 7, 20)


@pytest.mark.gpu
def test_gpu_grouped_scaler_equals_per_group():
    """(groups, samples, features): every group standardised with its own statistics in the same launches."""
    import torch
    from eeg2video_b200 import consumers
    rng = np.random.default_rng(21)
    x = torch.from_numpy((20 + 3 * rng.standard_normal((5, 333, 310))).astype(np.float32)).cuda()
    sc = consumers.StandardScaler().fit(x)
    y = sc.transform(x)
    assert tuple(sc.mean_.shape) == (5, 310) and tuple(y.shape) == (5, 333, 310)
    for g in range(5):
        one = consumers.StandardScaler().fit(x[g])
        assert torch.equal(one.mean_, sc.mean_[g]) and torch.equal(one.scale_, sc.scale_[g])
        assert torch.equal(one.transform(x[g]), y[g])
    with pytest.raises(ValueError, match="group axis"):
        sc.transform(x[0])


@pytest.mark.gpu
def test_gpu_graphed_build_replays_with_new_features(gold):
    """consumers.GraphedBuild: the six-kernel build captured as one CUDA graph; replay after refilling the input tensor
    in place must equal the kernel-by-kernel build of the new features, bit for bit."""
    import torch
    from eeg2video_b200 import _lib, consumers
    gt = gold["gt_label"]
    feat = torch.from_numpy(decode(gold["codes_1s"], np.float32)).cuda()
    chosen = list(range(1, 41))
    # the index table is uploaded BEFORE the capture (a host-to-device copy cannot be captured); the captured function
    # holds device work only: gather + mean over the windows, scaler fit, transform
    idx = torch.from_numpy(consumers.clip_index(range(6), gt, chosen)).cuda()

    def build(f):
        x = ops_mod.select_units(f.reshape(7 * 40 * 5, 2, -1), idx, True)
        return consumers.StandardScaler().fit_transform(x)
    from eeg2video_b200 import ops as ops_mod
    graphed = consumers.GraphedBuild(build, feat)
    want = consumers.semantic_predictor_inputs(feat, gt, chosen)[0]
    assert torch.equal(graphed.replay(), want)
    feat.add_(torch.randn(feat.shape, device=feat.device, generator=torch.Generator(feat.device).manual_seed(3)) * 500)
    # ^ new features in the same buffer (not an affine map of the old ones: standardising would undo that)
    want2 = consumers.semantic_predictor_inputs(feat, gt, chosen)[0]
    before = _lib.launch_count()
    got2 = graphed.replay()
    torch.cuda.synchronize()
    assert _lib.launch_count() == before                         # no launch goes through the library: one graph launch
    assert torch.equal(got2, want2) and not torch.equal(want2, want)
