// SPDX-License-Identifier: Apache-2.0
// Parameter block of the per-channel normalisation / activation kernels (rownorm.cu).
#pragma once

namespace wcn {

struct RowNormParams {
  const void* x;         // [n, ld_x] conv output
  const void* res;       // optional [n, ld_res] residual added before the activation
  const void* y_in;      // backward: forward output (ReLU mask), optional
  const void* dy;        // backward: upstream gradient [n, ld_dy]
  void* y;               // forward output / backward dx
  void* dres;            // backward: optional gradient of the residual (= masked dy)
  const float* scale;    // [c] forward: gamma * rstd; backward: gamma
  const float* shift;    // [c] forward: beta - mean * gamma * rstd
  const float* mean_rstd;  // [2c] saved batch mean, rstd (backward)
  // backward, optional: forward scale / shift; the ReLU mask is then recomputed as
  // x * mask_scale + mask_shift > 0 instead of being read from the saved output (one read less)
  const float* mask_scale;
  const float* mask_shift;
  double* sums;          // [2c] reduction target (stats: sum x, sum x^2; bwd: sum dz, sum dz*xhat)
  long long ld_x, ld_res, ld_y, ld_dy, ld_yin, ld_dres;  // row pitches in elements
  int n, c;
  int relu;
  int training;          // backward: 1 = batch statistics (full formula), 0 = dx = dz * scale
};

}  // namespace wcn
