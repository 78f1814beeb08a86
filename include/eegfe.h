/*
 * eegfe.h -- C ABI of libeegfe.so, the B200 (sm_100a) EEG feature front end.
 *
 * Drop-in boundary for the hot path of gaspachoo/EEG2Video's EEG_preprocessing package.  The reference has no
 * FFI layer of its own (it is pure Python); each entry point below therefore names the Python function whose
 * arithmetic it replaces, and eeg2video_b200/EEG_preprocessing/ mirrors those functions one-to-one on top of
 * this ABI (INTEGRATION.md shows the ctypes binding a maintainer of the reference would add).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers, borrowed for the duration of the call, never freed or retained.
 *   - Every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *     No hidden allocation, no host synchronisation.
 *   - Return value: 0 on success, a negative EEGFE_E* code for argument errors, or a positive cudaError_t.
 *     eegfe_error_string() turns either into text.
 *   - `status` (optional, may be NULL) points to one device int that the kernels OR flags into:
 *     EEGFE_STATUS_ZERO_POWER is set when some band has zero power, i.e. where the reference raises
 *     ValueError("math domain error") at DE_PSD.py:68.  The caller zeroes it and reads it back.
 *   - Feature layout is the reference's: for unit u (a clip or a pre-cut window group), window w, channel c,
 *     band b the value sits at  out[((u * n_windows + w) * n_ch + c) * 5 + b]  -- i.e. the
 *     (block, concept, repetition[, window], channel, band) arrays of the extract_DE_PSD_features_* drivers
 *     with the leading axes flattened.  Bands: delta, theta, alpha, beta, gamma (DE_PSD.py:28-29).
 *   - The fused kernels are built for 200 Hz and the drivers' three window lengths (FFT length 200, DE_PSD.py:27);
 *     eegfe_de_psd_generic covers every other (fre, time_window) the reference's DE_PSD accepts.
 */
#ifndef EEGFE_H_
#define EEGFE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEGFE_ABI_VERSION 1

/* analysis modes */
#define EEGFE_MODE_500MS 0 /* 7 windows of 100 samples, hop 50, Hann(100), zero-padded to 200 (1per500ms driver) */
#define EEGFE_MODE_1S 1    /* 2 windows of 200 samples, Hann(200)                               (1per1s script)  */
#define EEGFE_MODE_2S 2    /* 1 window: first 200 samples weighted by the head of Hann(400)     (1per2s driver)  */

/* argument errors */
#define EEGFE_EINVAL (-1)   /* bad mode / shape / null pointer                       */
#define EEGFE_ERANGE (-2)   /* block too short for the requested segments            */
#define EEGFE_EDTYPE (-3)   /* unsupported element type                              */

/* status flags */
#define EEGFE_STATUS_ZERO_POWER 1

/* element types for the byte-copy (segmentation) entry points */
#define EEGFE_DTYPE_F32 0
#define EEGFE_DTYPE_F64 1
#define EEGFE_DTYPE_F16 2
#define EEGFE_DTYPE_I16 3

int eegfe_abi_version(void);
const char* eegfe_error_string(int code);

/* Number of analysis windows per clip in `mode` (7, 2, 1), or EEGFE_EINVAL. */
int eegfe_windows_per_clip(int mode);

/*
 * Fused segmentation + DE/PSD straight from raw recordings.
 * Replaces: segment_raw_signals_200Hz.py:15-70 (extract_2s_segment index arithmetic, :58-65),
 *           segment_sliding_window.py:6-21, DE_PSD.py:8-71 and the loops of
 *           extract_DE_PSD_features_1per2s.py:16-28 / _1per1s.py:24-58 / _1per500ms.py:12-29.
 *
 * raw          float32 [n_blocks][n_ch][block_len] with strides (block_stride, ch_stride, 1) in elements;
 *              a "block" is one 8 min 40 s recording block of one subject (n_blocks = subjects * 7).
 * block_len    samples per row; must be >= 40 * 2600 = 104000 (else EEGFE_ERANGE, the reference's
 *              RuntimeError("Segment length mismatch"), segment_raw_signals_200Hz.py:68-69).
 * de, psd      float32 [n_blocks * 200][n_windows][n_ch][5]; clip index = (block * 40 + concept) * 5 + repetition.
 * No clip tensor is materialised: windows are cut by index arithmetic, clip (c, r) starting at sample
 * c * 2600 + 600 + r * 400.
 * Alignment: any (rows need only be float-aligned).  Rows that start on 16-byte boundaries are fetched by TMA bulk
 * copies; for other rows (block_len or a stride not a multiple of 4, a base pointer inside a buffer) the 16-byte
 * aligned span AROUND each row is fetched instead -- up to 12 bytes before its first and after its last sample -- and
 * read shifted.  Those bytes must be readable: true whenever the buffer is (part of) an allocation that starts and
 * ends on 16-byte boundaries (cudaMalloc, every framework allocator).
 */
int eegfe_de_psd_from_raw(const float* raw, int64_t n_blocks, int n_ch, int64_t block_len,
                          int64_t block_stride, int64_t ch_stride, int mode,
                          float* de, float* psd, int* status, void* stream);

/*
 * Same as eegfe_de_psd_from_raw for recordings whose 3 s hint periods have been dropped or moved: concept c of a
 * block starts at sample `first_offset + c * concept_stride` and holds its 5 repetitions back to back (5 x 400
 * samples).  eegfe_de_psd_from_raw is the special case first_offset = 600, concept_stride = 2600.  Used by the
 * host pipeline, which uploads only the 2000 live samples of every 2600 (one strided DMA per chunk, 23 % fewer
 * PCIe bytes): first_offset = 0, concept_stride = 2000, ch_stride = 80000.
 */
int eegfe_de_psd_from_concepts(const float* x, int64_t n_blocks, int n_ch, int64_t block_stride, int64_t ch_stride,
                               int64_t concept_stride, int64_t first_offset, int mode,
                               float* de, float* psd, int* status, void* stream);

/*
 * NEXT-ROW (BASELINE.json configs[4]): GLMNet input build in one pass over the raw recording.
 * The reference describes it (README.md:80-81, :88, :97-99: "Raw EEGs are normalized per channel using the training
 * split statistics"; model input contracts EEG-VP/models.py:119, :364) but its trainer / inference scripts are not in
 * the tree, so this row has no reference code to pin against (oracle/glmnet_inputs.py states the arithmetic).
 *
 * clips_norm   float32 [n_blocks * 200][n_ch][400] = (x - ch_mean[c]) * ch_scale[c]   (scale = 1/std; the subtraction
 *              comes first so that a large DC offset does not cancel digits); viewed as (N, 1, n_ch, 400) it is the
 *              input of glfnet / shallownet.
 * de, psd      float32 [n_blocks * 200][7][n_ch][5], the 500 ms features (identical to eegfe_de_psd_from_raw).
 * Requires 16-byte aligned rows (block_len, strides multiples of 4 samples) -> EEGFE_EINVAL otherwise.
 */
int eegfe_glmnet_inputs_from_raw(const float* raw, int64_t n_blocks, int n_ch, int64_t block_len, int64_t block_stride,
                                 int64_t ch_stride, const float* ch_scale, const float* ch_mean, float* clips_norm,
                                 float* de, float* psd, int* status, void* stream);

/*
 * Per-channel mean and population standard deviation of the clip samples (hint periods excluded) over the blocks
 * with block_mask[b] != 0 (NULL = all blocks): the "training split statistics" of the GLMNet raw branch.
 * workspace    double [n_blocks * n_ch * 2] (device scratch);  mean, std: double [n_ch] (device).
 * float64 accumulation in a fixed order (deterministic).
 */
int eegfe_channel_stats(const float* raw, int64_t n_blocks, int n_ch, int64_t block_len, int64_t block_stride,
                        int64_t ch_stride, const unsigned char* block_mask, double* workspace, double* mean,
                        double* std, void* stream);

/*
 * Strided host <-> device row copy on `stream` (cudaMemcpy2DAsync): `height` rows of `width` bytes, source /
 * destination pitches in bytes.  kind: 1 = host to device, 2 = device to host.  Plumbing for the host pipeline
 * (PyTorch has no strided pinned-memory DMA); no arithmetic.
 */
int eegfe_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width,
                       int64_t height, int kind, void* stream);

/*
 * DE/PSD of already segmented 2 s clips.
 * Replaces: extract_de_psd_raw (1per2s.py:16-28), the 1 s script body (1per1s.py:24-58) and -- fused with
 *           seg_sliding_window -- extract_de_psd_sw (1per500ms.py:12-29).
 * clips        float32 [n_clips][n_ch][400], contiguous.
 * de, psd      float32 [n_clips][n_windows][n_ch][5].
 */
int eegfe_de_psd_from_clips(const float* clips, int64_t n_clips, int n_ch, int mode,
                            float* de, float* psd, int* status, void* stream);

/*
 * DE/PSD of pre-cut windows: the DE_PSD(data, 200, time_window) call itself (DE_PSD.py:8-71) and
 * extract_de_psd_sw on a materialised (.., 7, 62, 100) tensor (1per500ms.py:24-25).
 * x            float32 [n_rows][win_len] with row stride `row_stride` elements (>= win_len);
 *              win_len = 100 (0.5 s), 200 (1 s) or 400 (2 s).
 * de, psd      float32 [n_rows][5].
 */
int eegfe_de_psd_windows(const float* x, int64_t n_rows, int win_len, int64_t row_stride,
                         float* de, float* psd, int* status, void* stream);

/*
 * Materialised segmentation (the .npy-producing scripts): byte-exact gathers, any 2/4/8-byte element type.
 * Replaces: segment_all_files (segment_raw_signals_200Hz.py:73-110)
 * raw -> clips [n_blocks * 200][n_ch][2 fs]; `fs` is the reference's sampling-rate argument (clip (c, r) starts
 * at c * 13 fs + 3 fs + r * 2 fs and block_len must be >= 40 * 13 fs).
 */
int eegfe_segment_clips(const void* raw, int dtype, int64_t n_blocks, int n_ch, int64_t block_len,
                        int64_t block_stride, int64_t ch_stride, int fs, void* clips, void* stream);

/*
 * Replaces: seg_sliding_window + np.save copy (segment_sliding_window.py:6-21, :55) for win 100 / hop 50:
 * clips [n_clips][n_ch][400] -> windows [n_clips][7][n_ch][100]
 */
int eegfe_sliding_windows(const void* clips, int dtype, int64_t n_clips, int n_ch, void* windows, void* stream);

/*
 * The same windows in either on-disk / in-memory layout the reference uses:
 *   EEGFE_WINDOWS_WINDOW_MAJOR  [n_clips][7][n_ch][100]   seg_sliding_window (segment_sliding_window.py:19)
 *   EEGFE_WINDOWS_LAST          [n_clips][n_ch][100][7]   the Seq2Seq trainer's inline loop, torch.stack(.., dim=-1)
 *                                                          (EEG2Video_New/Seq2Seq/my_autoregressive_transformer.py:309-314)
 * Byte-exact, any 2/4/8-byte element type.
 */
#define EEGFE_WINDOWS_WINDOW_MAJOR 0
#define EEGFE_WINDOWS_LAST 1
int eegfe_sliding_windows_layout(const void* clips, int dtype, int64_t n_clips, int n_ch, int layout, void* windows,
                                 void* stream);

/*
 * NEXT ROWS: the first thing every consumer of the feature files does.
 *
 * eegfe_select_units -- pick and re-order clips ("units" of [n_windows][n_cols] floats, n_cols = n_ch * 5), optionally
 * averaging the analysis windows.  Replaces the block selection + GT_label concept re-ordering + window mean of
 * train_semantic_predictor.py:87-95, :114 and eeg_text.py:115-125 (host computes src_index from the label table):
 *   reduce_windows == 0 : out[j][w][i] = feat[src_index[j]][w][i]                 out: [n_out][n_windows][n_cols]
 *   reduce_windows != 0 : out[j][i]    = mean_w feat[src_index[j]][w][i]          out: [n_out][n_cols]
 * src_index: device int32 [n_out], each in [0, n_units_in).
 */
int eegfe_select_units(const float* feat, int64_t n_units_in, int n_windows, int n_cols, const int* src_index,
                       int64_t n_out, int reduce_windows, float* out, void* stream);

/*
 * Column statistics and standardisation with sklearn.preprocessing.StandardScaler semantics (the reference's
 * `StandardScaler().fit(x); transform(x)` at EEG_VP_train_test.py:259-267, train_semantic_predictor.py:47-48,
 * eeg_text.py:142-144): mean and population variance per column in float64 (corrected two-pass algorithm, fixed
 * summation order), scale = sqrt(var) with 1 for near-constant columns (sklearn's _is_constant_feature bound);
 * transform = float32((double(x) - mean) / scale): the reference's call sites all run the scaler in float64 (float64
 * arrays, or torch tensors that scikit-learn converts to float64), so this is its result rounded once.  The third-party arithmetic is scikit-learn's (not pinned by the reference's requirements.txt;
 * restated from scikit-learn 1.9.0 as installed in the build container: preprocessing/_data.py, utils/extmath.py).
 * x: float32 [n_groups][n_rows][n_cols] with strides (group_stride, row_stride, 1) in elements; every group is an
 * independent matrix with its own statistics (one subject, one split).  mean, var, scale: device double
 * [n_groups][n_cols].  out: float32 [n_groups][n_rows][n_cols], contiguous.
 * workspace: device double [eegfe_column_stats_workspace(n_groups, n_rows, n_cols)].
 */
int64_t eegfe_column_stats_workspace(int64_t n_groups, int64_t n_rows, int n_cols);
int eegfe_column_stats(const float* x, int64_t n_groups, int64_t n_rows, int n_cols, int64_t row_stride,
                       int64_t group_stride, double* workspace, double* mean, double* var, double* scale, void* stream);
int eegfe_standardize(const float* x, int64_t n_groups, int64_t n_rows, int n_cols, int64_t row_stride,
                      int64_t group_stride, const double* mean, const double* scale, float* out, void* stream);

/*
 * Replaces: DE_PSD(data, fre, time_window) for ANY window length and sampling rate (DE_PSD.py:33-39, :49-58) -- the
 * general, slower path behind the three fused shapes.  The caller (eeg2video_b200/EEG_preprocessing/DE_PSD.py) evaluates
 * the reference's host expressions and passes their results:
 *   n_live      = min(L, 200), L = int(time_window * fre): samples of a row that enter fft(., 200) (:58)
 *   hann        : device float[n_live], 0.5 - 0.5 cos(2 pi (i + 1) / (L + 1)) (:51)
 *   band_lo/hi  : HOST int[5], inclusive bin range range(fStartNum - 1, fEndNum) with fNum = int(f / fre * 200) (:37-38,
 *                 :63); band_lo may be -1, which Python reads as the last element, bin 99; hi <= 99
 *   divisor     = hi - lo + 1 (:66)
 * x: float32 [n_rows] rows, row_stride elements apart.  de, psd: float32 [n_rows][5].
 */
int eegfe_de_psd_generic(const float* x, int64_t n_rows, int n_live, int64_t row_stride, const float* hann,
                         const int* band_lo, const int* band_hi, float* de, float* psd, int* status, void* stream);

/*
 * DE from PSD, elementwise: de[i] = log2(100 psd[i]) (DE_PSD.py:68) with the device expression of the feature kernels,
 * so the result is bit-identical to the DE those kernels wrote next to this PSD.  Used by the cohort gather: ranks send
 * PSD only and rank 0 rebuilds DE (half the NVLink bytes).  Zero power sets EEGFE_STATUS_ZERO_POWER.
 */
int eegfe_de_from_psd(const float* psd, int64_t n, float* de, int* status, void* stream);

/* Introspection used by bench.py / tests: kernel launch geometry chosen for `mode` on the current device. */
int eegfe_launch_geometry(int mode, int* grid, int* block, int* smem_bytes, int* rows_per_tile);

/* Number of kernels this library has launched since load (gpu_launches accounting in bench.py). */
int64_t eegfe_launch_count(void);

/* Cap the grid of the persistent feature kernels at `max_ctas` CTAs (0 = one per SM, the default); returns the previous
 * cap.  A CTA of these kernels fills its SM (registers, shared memory), so kernels of other streams -- NCCL's send /
 * receive kernels during the cohort gather -- only start when a CTA retires; leaving a few SMs free lets them overlap.
 * Process-wide. */
int eegfe_set_cta_limit(int max_ctas);

/* Tile loader of the 200-sample-row kernels (2 s mode, pre-cut 400-sample windows).  Default (0): one 1-D TMA
 * bulk copy per row.  1: clip-aligned tiles fetched by ONE TMA tensor copy each (cp.async.bulk.tensor, tensor map built
 * per launch, L2 promotion off) -- slower on B200 (6.0 vs 6.85 G channel-windows/s in 2 s mode), kept for traffic
 * measurements.  Process-wide; returns the previous setting.  eegfe_tma_launch_count(): launches that used it. */
int eegfe_set_tensor_loads(int on);
int64_t eegfe_tma_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EEGFE_H_ */
