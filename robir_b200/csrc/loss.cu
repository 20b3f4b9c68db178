// Fused PBR-stage loss: tone map + masked L1/L2 + latent-smooth L1 + KL sparsity + white-light
// regulariser, forward value AND every input gradient in one launch (the loss is a scalar, so the backward of the
// autograd node is a multiplication by the upstream gradient).  Replaces ~180 elementwise / reduction launches of
//   model/loss.py:61-125 (InvLoss.forward: get_rgb_loss, get_latent_smooth_loss, kl_divergence),
//   model/color_correction.py:31-59,131-133 (ACES hdr2ldr with the learnable exposure shift),
//   training/train_pbr.py:313-346 (white_loss, loss = rgb + kl + 0.1 smooth + white).
// One CTA of 1024 threads strides over the rays; sums are reduced in a fixed order (deterministic).
#include "common.cuh"
#include "loss_math.h"

namespace robir {

struct LossParams {
  int N;                 // rays
  int n_lat;             // rows of the latent (hit points; == N in the fixed-capacity mode)
  int M;                 // light SGs
  int l2;                // 0: L1 (hotdog.conf loss_type), 1: L2
  const float* sg_rgb; const float* indir_rgb;            // [N][3], row strides ld_sg / ld_ind (floats)
  int ld_sg, ld_ind;
  const float* gt;                                         // [N][3]   (ray order)
  const unsigned char* mask;                               // [N] network_object_mask & object_mask (ray order)
  const unsigned char* hit;                                // optional [N] network_object_mask (ray order), with order
  const long long* order;                                  // optional [N]: row i of the inputs is ray order[i] (the
                                                           // fixed-capacity path keeps hit rays compacted to the front)
  const float* adapt_illum;                                // [1] gamma.hdr_shift.adapt_illum
  const float* albedo; const float* albedo_r;              // [N][3], strides ld_alb / ld_albr
  int ld_alb, ld_albr;
  const float* rough; const float* rough_r;                // [N] (column 0), strides ld_r / ld_rr
  int ld_r, ld_rr;
  const float* z;                                          // [n_lat][32] latent pre-activation
  const unsigned char* z_valid;                            // [n_lat] or null (all valid)
  const float* lgt;                                        // [M][7]
  float w_rgb, w_kl, w_smooth, rho;
  float* losses;                                           // [5] total, sg_rgb_loss, kl, smooth, white
  float* g_pred;                                           // [N][3]  (gradient of sg_rgb and of indir_rgb)
  float* g_adapt;                                          // [1]
  float* g_albedo; float* g_albedo_r;                      // [N][3]
  float* g_rough; float* g_rough_r;                        // [N]
  float* g_z;                                              // [n_lat][32]
  float* g_lgt;                                            // [M][7]
};

constexpr int kLossThreads = 1024;
constexpr int kLossSums = 40;   // 0 rgb, 1 d rgb / d shift, 2 albedo L1, 3 rough L1, 4 n_valid, 5 white, 8..39 sigmoid(z) columns

__global__ void __launch_bounds__(kLossThreads, 1) pbr_loss_kernel(LossParams p) {
  __shared__ float s_part[32][kLossSums];
  __shared__ float s_tot[kLossSums];
  __shared__ float s_dkl[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float acc[kLossSums];
#pragma unroll
  for (int i = 0; i < kLossSums; ++i) acc[i] = 0.f;

  // exposure shift: clamp(clamp(10 a + 0.5, 0, 1), 1e-4, 1) ** 0.2   (color_correction.py:37-45)
  bool shift_live;
  const float shift = loss_shift(__ldg(p.adapt_illum), &shift_live);
  const float inv_s02 = powf(shift, -0.2f);
  const float inv_N = 1.f / (float)p.N;

  // ---- pass 1 over rays: rgb term (value + gradient), smooth terms (value + gradient)
  for (int i = tid; i < p.N; i += kLossThreads) {
    const long long ray = p.order ? p.order[i] : i;
    const float m = p.mask[ray] ? 1.f : 0.f;
    // compacted inputs: the twins of a ray that missed are both 1.0 in the reference's ray-order buffers
    const float hitrow = (p.order && p.hit) ? (p.hit[ray] ? 1.f : 0.f) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = p.sg_rgb[(size_t)i * p.ld_sg + c] + p.indir_rgb[(size_t)i * p.ld_ind + c];
      float per, d_pred, d_shift;
      loss_rgb_channel(x, p.gt[(size_t)ray * 3 + c], inv_s02, p.l2, m, &per, &d_pred, &d_shift);
      acc[0] += per;
      acc[1] += d_shift;                                                  // times d (s^-0.2) / d a below
      p.g_pred[(size_t)i * 3 + c] = p.w_rgb * inv_N * d_pred;
      const float da = hitrow * (p.albedo[(size_t)i * p.ld_alb + c] - p.albedo_r[(size_t)i * p.ld_albr + c]);
      acc[2] += fabsf(da);
      const float ga = p.w_smooth * loss_sgn(da) * inv_N * (1.f / 3.f);
      p.g_albedo[(size_t)i * 3 + c] = ga;
      p.g_albedo_r[(size_t)i * 3 + c] = -ga;
    }
    const float dr = hitrow * (p.rough[(size_t)i * p.ld_r] - p.rough_r[(size_t)i * p.ld_rr]);
    acc[3] += fabsf(dr);
    const float gr = p.w_smooth * 0.2f * loss_sgn(dr) * inv_N;
    p.g_rough[i] = gr;
    p.g_rough_r[i] = -gr;
  }
  // ---- latent: column sums of sigmoid(z) over the valid rows
  for (int i = tid; i < p.n_lat; i += kLossThreads) {
    if (p.z_valid != nullptr && !p.z_valid[i]) continue;
    acc[4] += 1.f;
    const float4* zr = reinterpret_cast<const float4*>(p.z + (size_t)i * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(zr + q);
      acc[8 + 4 * q] += loss_sigmoid(v.x);
      acc[9 + 4 * q] += loss_sigmoid(v.y);
      acc[10 + 4 * q] += loss_sigmoid(v.z);
      acc[11 + 4 * q] += loss_sigmoid(v.w);
    }
  }
  // ---- white-light regulariser: var_c(|mu| / (||mu|| + 1e-4)) averaged over the lobes, * 0.01  (train_pbr.py:313-317)
  for (int i = tid; i < p.M; i += kLossThreads) {
    const float* r = p.lgt + (size_t)i * 7;
    const float x[3] = {r[4], r[5], r[6]};
    float gx[3];
    acc[5] += loss_white_lobe(x, gx);
    const float scale = 0.01f / (float)p.M;
    float* g = p.g_lgt + (size_t)i * 7;
    g[0] = g[1] = g[2] = g[3] = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) g[4 + j] = scale * gx[j];
  }
  // ---- block reduction (fixed order)
#pragma unroll
  for (int i = 0; i < kLossSums; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) s_part[warp][i] = v;
  }
  __syncthreads();
  if (tid < kLossSums) {
    float t = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) t += s_part[w][tid];
    s_tot[tid] = t;
  }
  __syncthreads();
  const float n_valid = fmaxf(s_tot[4], 1.f);
  if (tid < 32) {
    const float rh = s_tot[8 + tid] / n_valid;
    float d_rh;
    s_part[0][tid] = loss_kl_column(p.rho, rh, &d_rh);                    // reuse as scratch (all reads are done)
    s_dkl[tid] = d_rh * (1.f / 32.f);
  }
  __syncthreads();
  if (tid == 0) {
    float kl = 0.f;
    for (int j = 0; j < 32; ++j) kl += s_part[0][j];
    kl *= (1.f / 32.f);
    const float rgb = s_tot[0] * inv_N;
    const float smooth = s_tot[2] * inv_N * (1.f / 3.f) + 0.2f * s_tot[3] * inv_N;
    const float white = 0.01f * s_tot[5] / (float)p.M;
    p.losses[0] = p.w_rgb * rgb + p.w_kl * kl + p.w_smooth * smooth + white;
    p.losses[1] = rgb;
    p.losses[2] = kl;
    p.losses[3] = smooth;
    p.losses[4] = white;
    // d / d adapt_illum through ldr = aces / s^0.2
    p.g_adapt[0] = shift_live ? p.w_rgb * inv_N * s_tot[1] * (-0.2f) * powf(shift, -1.2f) * 10.f : 0.f;
  }
  // ---- pass 2: KL gradient w.r.t. the latent
  for (int i = tid; i < p.n_lat; i += kLossThreads) {
    const bool ok = p.z_valid == nullptr || p.z_valid[i];
    const float4* zr = reinterpret_cast<const float4*>(p.z + (size_t)i * 32);
    float4* gz = reinterpret_cast<float4*>(p.g_z + (size_t)i * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float4 v = __ldg(zr + q);
        const float k = p.w_kl / n_valid;
        const float s0 = loss_sigmoid(v.x), s1 = loss_sigmoid(v.y);
        const float s2 = loss_sigmoid(v.z), s3 = loss_sigmoid(v.w);
        o.x = k * s_dkl[4 * q] * s0 * (1.f - s0);
        o.y = k * s_dkl[4 * q + 1] * s1 * (1.f - s1);
        o.z = k * s_dkl[4 * q + 2] * s2 * (1.f - s2);
        o.w = k * s_dkl[4 * q + 3] * s3 * (1.f - s3);
      }
      gz[q] = o;
    }
  }
}

}  // namespace robir

using namespace robir;

extern "C" {

// Value + gradients of the PBR-stage training loss.  All buffers are caller-owned; z rows are 32 floats (latent_dim).
int robir_pbr_loss(const LossParams* p, void* stream) {
  RB_REQUIRE(p->N > 0 && p->M > 0, "pbr_loss: empty batch");
  pbr_loss_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
