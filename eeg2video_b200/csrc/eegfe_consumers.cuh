// NEXT ROWS: what every consumer of the feature files does first --
// pick blocks, re-order the 40 concepts of each block by the label table, average the analysis windows, flatten
// (channel, band) to 310 columns, standardise the columns (sklearn StandardScaler).  Reference call sites:
//   EEG-VP/EEG_VP_train_test.py:232, :254-267          (rearrange, reshape to 62*5, StandardScaler per split)
//   EEG2Video_New/Generation/models/train_semantic_predictor.py:47-48, :87-95, :114   (concept order, mean over the two
//                                                        1 s windows, 310 columns, StandardScaler)
//   EEG2Video_New/Semantic/eeg_text.py:115-125, :142-144 (same on the 2 s features)
// Small HBM-bound kernels; statistics in float64 with a fixed summation order (deterministic).
#pragma once

namespace eegfe {

// out[j][i] for output unit j (a clip) and column i < cols = n_ch * 5:
//   keep windows : out[(j * W + w) * cols + i] = feat[(src[j] * W + w) * cols + i]
//   mean windows : out[j * cols + i] = (sum_w feat[(src[j] * W + w) * cols + i]) / W     (summed in window order)
__global__ void __launch_bounds__(512) select_units_kernel(const float* __restrict__ feat, const int* __restrict__ src,
                                                            long long n_out, int n_windows, int cols, int reduce,
                                                            float* __restrict__ out)
{
  // one output unit per block step, threads along the unit's columns: no per-element division, coalesced both ways
  const int per_unit = reduce ? cols : n_windows * cols;
  const float inv_w = 1.0f / static_cast<float>(n_windows);
  for (long long j = blockIdx.x; j < n_out; j += gridDim.x) {
    const float* unit = feat + static_cast<long long>(src[j]) * n_windows * cols;
    float* o = out + j * per_unit;
    for (int i = threadIdx.x; i < per_unit; i += blockDim.x) {
      if (reduce) {
        float acc = unit[i];
        for (int w = 1; w < n_windows; ++w) acc = __fadd_rn(acc, unit[static_cast<long long>(w) * cols + i]);
        o[i] = n_windows == 1 ? acc : (n_windows == 2 ? acc * 0.5f : acc * inv_w);
      } else {
        o[i] = unit[i];
      }
    }
  }
}

constexpr int kStatRowsPerBlock = 64;

// PASS 0: partial[b][c] = sum over the rows of chunk b of x[r][c].
// PASS 1: partial[b][c] = sum of (x - mean)^2 and partial[n_chunks + b][c] = sum of (x - mean) (the correction term of
//         the corrected two-pass algorithm, sklearn/utils/extmath.py _incremental_mean_and_var).  float64, rows in order.
template <int PASS>
__global__ void __launch_bounds__(512) column_partial_kernel(const float* __restrict__ x, long long n_rows, int n_cols,
                                                              long long row_stride, long long group_stride,
                                                              const double* __restrict__ shift,
                                                              double* __restrict__ partial)
{
  // blockIdx.y = group (an independent matrix with its own statistics, e.g. one subject)
  x += blockIdx.y * group_stride;
  partial += static_cast<long long>(blockIdx.y) * 2 * gridDim.x * n_cols;
  if (PASS == 1) shift += static_cast<long long>(blockIdx.y) * n_cols;
  const long long r0 = static_cast<long long>(blockIdx.x) * kStatRowsPerBlock;
  const long long r1 = (r0 + kStatRowsPerBlock < n_rows) ? r0 + kStatRowsPerBlock : n_rows;
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    const double m = PASS == 1 ? shift[c] : 0.0;
    double acc = 0.0, lin = 0.0;
    // eight loads in flight per thread, summed in row order (the order is part of the result: deterministic)
    for (long long rb = r0; rb < r1; rb += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = rb + k < r1 ? x[(rb + k) * row_stride + c] : 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (rb + k < r1) {
          const double d = static_cast<double>(v[k]) - m;
          if (PASS == 1) {
            acc += d * d;
            lin += d;
          } else {
            acc += d;
          }
        }
      }
    }
    partial[static_cast<long long>(blockIdx.x) * n_cols + c] = acc;
    if (PASS == 1) partial[(static_cast<long long>(gridDim.x) + blockIdx.x) * n_cols + c] = lin;
  }
}

// PASS 0: mean[c] = (sum_b partial[b][c]) / n.
// PASS 1: var[c] = (sum of squares - correction^2 / n) / n, scale[c] = sqrt(var[c]), with sklearn's rule for constant
// columns (sklearn/preprocessing/_data.py _is_constant_feature): scale 1 where var <= n eps var + (n mean eps)^2.
template <int PASS>
__global__ void column_finish_kernel(const double* __restrict__ partial, int n_chunks, long long n_rows, int n_cols,
                                     const double* __restrict__ mean_in, double* __restrict__ out_a,
                                     double* __restrict__ out_b)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  partial += static_cast<long long>(blockIdx.y) * 2 * n_chunks * n_cols;
  const long long g = static_cast<long long>(blockIdx.y) * n_cols;
  if (PASS == 1) mean_in += g;
  out_a += g;
  if (PASS == 1) out_b += g;
  double acc = 0.0;
  for (int b = 0; b < n_chunks; ++b) acc += partial[static_cast<long long>(b) * n_cols + c];
  const double n = static_cast<double>(n_rows);
  if (PASS == 0) {
    out_a[c] = n_rows > 0 ? acc / n : 0.0;
  } else {
    double lin = 0.0;
    for (int b = 0; b < n_chunks; ++b) lin += partial[static_cast<long long>(n_chunks + b) * n_cols + c];
    const double var = n_rows > 0 ? (acc - lin * lin / n) / n : 0.0;
    const double m = mean_in[c];
    const double eps = 2.220446049250313e-16;
    const bool constant = var <= n * eps * var + (n * m * eps) * (n * m * eps);
    out_a[c] = var;
    out_b[c] = constant ? 1.0 : sqrt(var);
  }
}

// out = float32((double(x) - mean) / scale).  Every call site of the reference hands StandardScaler either a float64
// numpy array (EEG_VP_train_test.py:259-267) or a torch tensor, which scikit-learn converts to float64 before
// `X -= mean_; X /= scale_` (train_semantic_predictor.py:47-48, eeg_text.py:142-144): the reference's result is the
// float64 quotient, and this is its correctly rounded float32 value.
__global__ void __launch_bounds__(512) standardize_kernel(const float* __restrict__ x, long long n_rows, int n_cols,
                                                           long long row_stride, long long group_stride,
                                                           const double* __restrict__ mean,
                                                           const double* __restrict__ scale, float* __restrict__ out)
{
  x += blockIdx.y * group_stride;
  mean += static_cast<long long>(blockIdx.y) * n_cols;
  scale += static_cast<long long>(blockIdx.y) * n_cols;
  out += static_cast<long long>(blockIdx.y) * n_rows * n_cols;
  // threads along the columns (a thread keeps its column's mean / scale in registers), block steps over the rows:
  // coalesced, no 64-bit division per element
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    const double m = mean[c], sc = scale[c];
    for (long long r = blockIdx.x; r < n_rows; r += gridDim.x)
      out[r * n_cols + c] = static_cast<float>((static_cast<double>(x[r * row_stride + c]) - m) / sc);
  }
}

}  // namespace eegfe
