// TEST INFRASTRUCTURE ONLY -- never linked into, or imported by, the product (eeg2video_b200/).
//
// Compiles the kernel's arithmetic core (eeg2video_b200/csrc/bandpower.cuh, scalar backend, explicitly rounded
// fp32) for the HOST so that the CPU-only test tier can check the exact butterfly order, index maps and
// constants of the CUDA kernel against the oracle without a GPU.  The band energies it returns are bit-identical
// to what the device code computes (same operations, same order, IEEE fp32 add / mul / fma).
#include <cstdint>
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
#include "../../eeg2video_b200/csrc/bandpower.cuh"

extern "C" {

// x: n_windows rows of `len` float32 samples (len = 100: 500 ms, 200: 1 s, 400: 2 s), contiguous.
// energy: n_windows x 5 float32, E_b = sum_{k in band b} |X[k]|^2.
int hostemu_band_energy(const float* x, int64_t n_windows, int len, float* energy)
{
  alignas(16) float buf[200];
  for (int64_t w = 0; w < n_windows; ++w) {
    const float* row = x + w * len;
    float e[5];
    if (len == 100) {
      for (int i = 0; i < 100; ++i) buf[i] = row[i];
      eegfe::window_band_energy<4, eegfe::kHannHalfSec, 2>(buf, e);
    } else if (len == 200) {
      for (int i = 0; i < 200; ++i) buf[i] = row[i];
      eegfe::window_band_energy<8, eegfe::kHannOneSec, 4>(buf, e);
    } else if (len == 400) {
      for (int i = 0; i < 200; ++i) buf[i] = row[i];
      eegfe::window_band_energy<8, eegfe::kHannTwoSec, 4>(buf, e);
    } else {
      return 1;
    }
    for (int b = 0; b < 5; ++b) energy[w * 5 + b] = e[b];
  }
  return 0;
}

// The shifted-span form of the kernels (rows that are not 16-byte aligned): the window starts `shift` floats (0..3) into
// a 16-byte aligned row and is read with scalar loads (vec = 1) or, for even shifts, 64-bit loads (vec = 2).
int hostemu_band_energy_shifted(const float* x, int64_t n_windows, int len, int shift, int vec, float* energy)
{
  if (shift < 0 || shift > 3 || (vec != 1 && vec != 2) || (vec == 2 && (shift & 1))) return 1;
  alignas(16) float buf[208];
  for (int64_t w = 0; w < n_windows; ++w) {
    const float* row = x + w * len;
    float e[5];
    for (int i = 0; i < 208; ++i) buf[i] = -12345.0f;                 // anything outside the window must not matter
    const int live = len == 100 ? 100 : 200;
    for (int i = 0; i < live; ++i) buf[shift + i] = row[i];
    const float* win = buf + shift;
    if (len == 100) {
      if (vec == 1) eegfe::window_band_energy<4, eegfe::kHannHalfSec, 1>(win, e);
      else eegfe::window_band_energy<4, eegfe::kHannHalfSec, 2>(win, e);
    } else if (len == 200) {
      if (vec == 1) eegfe::window_band_energy<8, eegfe::kHannOneSec, 1>(win, e);
      else eegfe::window_band_energy<8, eegfe::kHannOneSec, 2>(win, e);
    } else if (len == 400) {
      if (vec == 1) eegfe::window_band_energy<8, eegfe::kHannTwoSec, 1>(win, e);
      else eegfe::window_band_energy<8, eegfe::kHannTwoSec, 2>(win, e);
    } else {
      return 1;
    }
    for (int b = 0; b < 5; ++b) energy[w * 5 + b] = e[b];
  }
  return 0;
}

}  // extern "C"
