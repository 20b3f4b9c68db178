// Stage-1 NeuS volume renderer, the per-ray parts around the SDF / colour networks:
// hierarchical importance sampling (up_sample + sample_pdf + cat_z_vals) and the alpha compositing of render_core.
// Reference: neus/volume_render/sdf_render.py -- sample_pdf :5-35 (det=True), up_sample :38-82, cat_z_vals :85-99,
// render_core :141-233 (n_outside = 0: no background model), render_neus :236-348.  Forward (evaluation) path; the
// network evaluations between these kernels are robir_sdf_eval (value + normal + features) and the fused colour chain.
// One warp per ray, the sample axis (<= 256 depths) lives in shared memory; the scans are sequential in lane 0 like
// torch's CPU cumsum / cumprod (fp32, same order of operations as the reference's CPU path).
#include "common.cuh"

namespace robir {

constexpr int kNeusMaxSamples = 256;
constexpr int kNeusWarps = 4;

__device__ __forceinline__ float neus_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// torch.linspace(lo, hi, steps)[p] in fp32 (ATen: symmetric evaluation around the midpoint)
__device__ __forceinline__ float neus_linspace(float lo, float hi, int steps, int p) {
  const float step = (hi - lo) / (float)(steps - 1);
  return (p < steps / 2) ? (lo + step * (float)p) : (hi - step * (float)(steps - p - 1));
}

// new_z [B][n_imp] = sample_pdf(z, section weights of the current SDF samples at a fixed inv_s)  (sdf_render.py:38-82)
__global__ void __launch_bounds__(32 * kNeusWarps) neus_upsample_kernel(
    int B, int n, int n_imp, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
    const float* __restrict__ z, const float* __restrict__ sdf, float inv_s, float radius, float* __restrict__ new_z) {
  __shared__ float s_z[kNeusWarps][kNeusMaxSamples], s_f[kNeusWarps][kNeusMaxSamples], s_w[kNeusWarps][kNeusMaxSamples],
      s_cdf[kNeusWarps][kNeusMaxSamples + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kNeusWarps + warp;
  if (ray >= B) return;
  float* zz = s_z[warp]; float* ff = s_f[warp]; float* ww = s_w[warp]; float* cdf = s_cdf[warp];
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  for (int i = lane; i < n; i += 32) { zz[i] = z[(size_t)ray * n + i]; ff[i] = sdf[(size_t)ray * n + i]; }
  __syncwarp();
  // per-section slope, clipped to the non-positive "entering" side and zeroed outside the sphere
  for (int i = lane; i < n - 1; i += 32) {
    const float cosv = (ff[i + 1] - ff[i]) / (zz[i + 1] - zz[i] + 1e-5f);
    const float prev = i == 0 ? 0.f : (ff[i] - ff[i - 1]) / (zz[i] - zz[i - 1] + 1e-5f);
    const float px0 = ox + dx * zz[i], py0 = oy + dy * zz[i], pz0 = oz + dz * zz[i];
    const float px1 = ox + dx * zz[i + 1], py1 = oy + dy * zz[i + 1], pz1 = oz + dz * zz[i + 1];
    const bool inside = sqrtf(px0 * px0 + py0 * py0 + pz0 * pz0) < radius ||
                        sqrtf(px1 * px1 + py1 * py1 + pz1 * pz1) < radius;
    float c = fminf(prev, cosv);
    c = fminf(fmaxf(c, -1e3f), 0.f) * (inside ? 1.f : 0.f);
    const float mid = (ff[i] + ff[i + 1]) * 0.5f, dist = zz[i + 1] - zz[i];
    const float pc = neus_sigmoid((mid - c * dist * 0.5f) * inv_s), nc = neus_sigmoid((mid + c * dist * 0.5f) * inv_s);
    ww[i] = (pc - nc + 1e-5f) / (pc + 1e-5f);               // alpha
  }
  __syncwarp();
  if (lane == 0) {
    float T = 1.f, sum = 0.f;
    for (int i = 0; i < n - 1; ++i) {                        // weights = alpha * exclusive cumprod(1 - alpha + 1e-7)
      const float a = ww[i];
      const float w = a * T + 1e-5f;                         // (+ 1e-5: sample_pdf :7)
      T *= (1.f - a + 1e-7f);
      ww[i] = w;
      sum += w;
    }
    float c = 0.f;
    cdf[0] = 0.f;
    for (int i = 0; i < n - 1; ++i) { c += ww[i] / sum; cdf[i + 1] = c; }
  }
  __syncwarp();
  for (int k = lane; k < n_imp; k += 32) {
    const float u = neus_linspace(0.5f / (float)n_imp, 1.f - 0.5f / (float)n_imp, n_imp, k);
    int lo = 0, hi = n;                                      // searchsorted(cdf[0..n), u, right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = lo - 1 < 0 ? 0 : lo - 1, above = lo > n - 1 ? n - 1 : lo;
    float denom = cdf[above] - cdf[below];
    denom = denom < 1e-5f ? 1.f : denom;
    new_z[(size_t)ray * n_imp + k] = zz[below] + (u - cdf[below]) / denom * (zz[above] - zz[below]);
  }
}

// torch.sort(cat([z, new_z])) + the same gather of the SDF values (cat_z_vals, :85-99).  Both lists are ascending;
// ties keep the old depth first (stable sort of the concatenation).  sdf / new_sdf / out_sdf may be null.
__global__ void neus_merge_kernel(int B, int n, int m, const float* __restrict__ z, const float* __restrict__ sdf,
                                  const float* __restrict__ new_z, const float* __restrict__ new_sdf,
                                  float* __restrict__ out_z, float* __restrict__ out_sdf) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= B) return;
  const float* a = z + (size_t)ray * n;
  const float* b = new_z + (size_t)ray * m;
  int i = 0, j = 0;
  for (int k = 0; k < n + m; ++k) {
    const bool take_a = j >= m || (i < n && a[i] <= b[j]);
    out_z[(size_t)ray * (n + m) + k] = take_a ? a[i] : b[j];
    if (out_sdf != nullptr)
      out_sdf[(size_t)ray * (n + m) + k] = take_a ? sdf[(size_t)ray * n + i] : new_sdf[(size_t)ray * m + j];
    if (take_a) ++i; else ++j;
  }
}

// section midpoints of render_core (:150-156): mid_z [B][n], pts [B*n][3]
__global__ void neus_midpoints_kernel(int B, int n, float sample_dist, const float* __restrict__ rays_o,
                                      const float* __restrict__ rays_d, const float* __restrict__ z,
                                      float* __restrict__ mid_z, float* __restrict__ pts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * n) return;
  const int ray = idx / n, i = idx % n;
  const float z0 = z[idx];
  const float dist = i + 1 < n ? z[idx + 1] - z0 : sample_dist;
  const float mz = z0 + dist * 0.5f;
  mid_z[idx] = mz;
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[3 * (size_t)idx + c] = rays_o[3 * ray + c] + rays_d[3 * ray + c] * mz;
}

// alpha compositing of render_core (:158-233) + the per-ray reductions render_neus adds (:333-341).
// eik_acc [2]: sum(relax_inside * (|grad| - 1)^2), sum(relax_inside)  (atomic, zero-initialised)
__global__ void __launch_bounds__(32 * kNeusWarps) neus_composite_kernel(
    int B, int n, float sample_dist, float inv_s, float cos_anneal, float radius, int white_bkgd,
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z,
    const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ color,
    const float* __restrict__ near, const float* __restrict__ far, float* __restrict__ rgb, float* __restrict__ weights,
    float* __restrict__ acc, float* __restrict__ dist_out, float* __restrict__ eik_acc) {
  __shared__ float s_a[kNeusWarps][kNeusMaxSamples];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kNeusWarps + warp;
  if (ray >= B) return;
  float* al = s_a[warp];
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  float e_num = 0.f, e_den = 0.f;
  for (int i = lane; i < n; i += 32) {
    const size_t s = (size_t)ray * n + i;
    const float z0 = z[s];
    const float dist = i + 1 < n ? z[s + 1] - z0 : sample_dist;
    const float mz = z0 + dist * 0.5f;
    const float px = ox + dx * mz, py = oy + dy * mz, pz = oz + dz * mz;
    const float gx = grad[3 * s], gy = grad[3 * s + 1], gz = grad[3 * s + 2];
    const float true_cos = dx * gx + dy * gy + dz * gz;
    const float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.f - cos_anneal) + fmaxf(-true_cos, 0.f) * cos_anneal);
    const float f = sdf[s];
    const float pc = neus_sigmoid((f - iter_cos * dist * 0.5f) * inv_s), nc = neus_sigmoid((f + iter_cos * dist * 0.5f) * inv_s);
    float a = (pc - nc + 1e-5f) / (pc + 1e-5f);
    a = fminf(fmaxf(a, 0.f), 1.f);
    const float nrm = sqrtf(px * px + py * py + pz * pz);
    al[i] = nrm < radius ? a : 0.f;
    if (nrm < radius * 1.2f) {
      const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.f;
      e_num += e * e;
      e_den += 1.f;
    }
  }
  __syncwarp();
  if (lane == 0) {                                           // weights = alpha * exclusive cumprod(1 - alpha + 1e-7)
    float T = 1.f;
    for (int i = 0; i < n; ++i) {
      const float a = al[i];
      al[i] = a * T;
      T *= (1.f - a + 1e-7f);
    }
  }
  __syncwarp();
  float r = 0.f, g = 0.f, b = 0.f, wsum = 0.f, wz = 0.f;
  for (int i = lane; i < n; i += 32) {
    const size_t s = (size_t)ray * n + i;
    const float w = al[i];
    weights[s] = w;
    r += color[3 * s] * w; g += color[3 * s + 1] * w; b += color[3 * s + 2] * w;
    wsum += w;
    if (i < 128) {                                           // render_neus :336 uses weights[..., :n_samples + n_importance]
      const float z0 = z[s];
      const float dist = i + 1 < n ? z[s + 1] - z0 : sample_dist;
      wz += w * (z0 + dist * 0.5f);
    }
  }
  r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); wsum = warp_sum(wsum); wz = warp_sum(wz);
  e_num = warp_sum(e_num); e_den = warp_sum(e_den);
  if (lane == 0) {
    const float bg = white_bkgd ? (1.f - wsum) : 0.f;
    rgb[3 * ray] = r + bg; rgb[3 * ray + 1] = g + bg; rgb[3 * ray + 2] = b + bg;
    acc[ray] = wsum;
    float d = wz / wsum;
    if (d != d) d = INFINITY;                                // nan_to_num(distance, inf)
    d = fminf(fmaxf(d, near[ray]), far[ray]);
    dist_out[ray] = d;
    atomicAdd(eik_acc, e_num);
    atomicAdd(eik_acc + 1, e_den);
  }
}

}  // namespace robir

using namespace robir;

extern "C" {

int robir_neus_upsample(int B, int n, int n_imp, const float* rays_o, const float* rays_d, const float* z,
                        const float* sdf, float inv_s, float radius, float* new_z, void* stream) {
  if (B == 0) return 0;
  RB_REQUIRE(n >= 2 && n <= kNeusMaxSamples && n_imp >= 1, "neus_upsample: 2 <= n <= 256 depths per ray");
  neus_upsample_kernel<<<(B + kNeusWarps - 1) / kNeusWarps, 32 * kNeusWarps, 0, (cudaStream_t)stream>>>(
      B, n, n_imp, rays_o, rays_d, z, sdf, inv_s, radius, new_z);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_neus_merge(int B, int n, int m, const float* z, const float* sdf, const float* new_z, const float* new_sdf,
                     float* out_z, float* out_sdf, void* stream) {
  if (B == 0) return 0;
  RB_REQUIRE((sdf == nullptr) == (out_sdf == nullptr) || out_sdf == nullptr, "neus_merge: out_sdf needs sdf + new_sdf");
  neus_merge_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, n, m, z, sdf, new_z, new_sdf, out_z, out_sdf);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_neus_midpoints(int B, int n, float sample_dist, const float* rays_o, const float* rays_d, const float* z,
                         float* mid_z, float* pts, void* stream) {
  if (B == 0) return 0;
  neus_midpoints_kernel<<<(B * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, n, sample_dist, rays_o, rays_d, z,
                                                                               mid_z, pts);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_neus_composite(int B, int n, float sample_dist, float inv_s, float cos_anneal, float radius, int white_bkgd,
                         const float* rays_o, const float* rays_d, const float* z, const float* sdf, const float* grad,
                         const float* color, const float* near, const float* far, float* rgb, float* weights,
                         float* acc, float* dist, float* eik_acc, void* stream) {
  if (B == 0) return 0;
  RB_REQUIRE(n >= 1 && n <= kNeusMaxSamples, "neus_composite: at most 256 depths per ray");
  neus_composite_kernel<<<(B + kNeusWarps - 1) / kNeusWarps, 32 * kNeusWarps, 0, (cudaStream_t)stream>>>(
      B, n, sample_dist, inv_s, cos_anneal, radius, white_bkgd, rays_o, rays_d, z, sdf, grad, color, near, far, rgb,
      weights, acc, dist, eik_acc);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
