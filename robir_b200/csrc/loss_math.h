// Per-element math of the fused PBR loss (loss.cu), host+device so that tests/hostcheck can check it against torch
// autograd on machines without a GPU.  Reference behaviour restated from model/loss.py:61-125 (InvLoss),
// model/color_correction.py:31-59 (ACES hdr2ldr with the learnable exposure), training/train_pbr.py:313-346 (white_loss).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RB_LHD __host__ __device__ __forceinline__
#else
#define RB_LHD inline
#endif

namespace robir {

RB_LHD float loss_sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// exposure shift s = clamp(clamp(10 a + 0.5, 0, 1), 1e-4, 1); live = d s / d a != 0 (torch.clamp: inclusive bounds)
RB_LHD float loss_shift(float a, bool* live) {
  const float raw = 10.f * a + 0.5f;
  *live = raw >= 1e-4f && raw <= 1.f;
  return fminf(fmaxf(fminf(fmaxf(raw, 0.f), 1.f), 1e-4f), 1.f);
}

// one colour channel of the rgb term: ldr = aces(x) / s^0.2, per = |ldr - gt| (or squared) * m.
//   value    : per
//   d_pred   : d per / d x           (caller multiplies by w_rgb / N)
//   d_shift  : d per / d (s^-0.2) = dper * aces(x)   (caller multiplies by d (s^-0.2) / d a)
RB_LHD void loss_rgb_channel(float x, float gt, float inv_s02, int l2, float m, float* value, float* d_pred,
                             float* d_shift) {
  const float num = x * (2.51f * x + 0.03f), den = x * (2.43f * x + 0.59f) + 0.14f;
  const float ac = num / den;
  const float dac = ((5.02f * x + 0.03f) * den - num * (4.86f * x + 0.59f)) / (den * den);
  const float diff = ac * inv_s02 - gt;
  const float dper = (l2 ? 2.f * diff : loss_sgn(diff)) * m;
  *value = (l2 ? diff * diff : fabsf(diff)) * m;
  *d_pred = dper * dac * inv_s02;
  *d_shift = dper * ac;
}

// white-light regulariser of one lobe: var_c(|mu| / (||mu|| + 1e-4)) (unbiased over the 3 channels) and its gradient
RB_LHD float loss_white_lobe(const float (&x)[3], float (&g)[3]) {
  const float c[3] = {fabsf(x[0]), fabsf(x[1]), fabsf(x[2])};
  const float nrm = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  const float mu = nrm + 1e-4f;
  const float u[3] = {c[0] / mu, c[1] / mu, c[2] / mu};
  const float ub = (u[0] + u[1] + u[2]) * (1.f / 3.f);
  const float d[3] = {u[0] - ub, u[1] - ub, u[2] - ub};
  // d var / d u_k = d_k;  d u_k / d c_j = delta_kj / mu - c_k c_j / (mu^2 nrm)
  const float dc = d[0] * c[0] + d[1] * c[1] + d[2] * c[2];
  for (int j = 0; j < 3; ++j) {
    const float gcj = d[j] / mu - (nrm > 0.f ? dc * c[j] / (mu * mu * nrm) : 0.f);
    g[j] = gcj * loss_sgn(x[j]);
  }
  return 0.5f * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

// KL sparsity of one latent column with mean activation rh: term and d term / d rh
RB_LHD float loss_kl_column(float rho, float rh, float* d_rh) {
  *d_rh = -rho / (rh + 1e-4f) + (1.f - rho) / (1.f - rh + 1e-4f);
  return rho * logf(rho / (rh + 1e-4f)) + (1.f - rho) * logf((1.f - rho) / (1.f - rh + 1e-4f));
}

RB_LHD float loss_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace robir
