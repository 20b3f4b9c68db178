// Fused spherical-Gaussian render at the hit points: direct lights (M shared lobes, diffuse +
// specular, with per-lobe light visibility and per-point BRDF visibility) and indirect lights (Mi per-point lobes,
// specular only; diffuse = indir_integral * albedo / pi), forward and backward in one launch each.
// Reference: model/sg_render.py:304-337 (render_with_all_sg), :343-565 (render_with_sg), single view, metallic=None,
// fun_spec=False, diffuse_vis=None.  Math lives in sg_math.h; the backward uses forward-mode duals per (point, lobe).
#include "common.cuh"
#include "sg_math.h"

namespace robir {

struct SgParams {
  int n, M, Mi;
  int lin_diff;
  const float* normal;        // [n][3]
  const float* view;          // [n][3] unit, towards camera
  const float* rough;         // [n]
  const float* albedo;        // [n][3]
  const float* spec_refl;     // [1]  (already abs()'ed by the caller, train_pbr.py:377)
  const float* lgt;           // [M][7] raw direct SGs
  const float* ind_lgt;       // [n][Mi][7] raw indirect SGs (may be null when Mi == 0)
  const float* light_vis;     // [n][M]
  const float* bv_dir;        // [n]
  const float* bv_ind;        // [n]
  const float* ind_integral;  // [n][3] (already * 2 pi, train_pbr.py:365)
  // forward outputs
  float* sg_rgb; float* sg_spec; float* sg_diff; float* vis_shadow;
  float* ind_rgb; float* ind_spec; float* ind_diff;
  float* pre;                 // [n][9] pre-clamp sums: direct spec(3), direct diff(3), indirect spec(3)
  // backward inputs (null = zero)
  const float* g_sg_rgb; const float* g_sg_spec; const float* g_sg_diff;
  const float* g_ind_rgb; const float* g_ind_spec; const float* g_ind_diff;
  // backward outputs
  float* g_lgt;               // [M][7] atomically accumulated (zero-init by caller)
  float* g_ind_lgt;           // [n][Mi][7]
  float* g_light_vis;         // [n][M]
  float* g_bv_dir; float* g_bv_ind; float* g_rough;   // [n]
  float* g_albedo;            // [n][3]
  float* g_spec_refl;         // [1] atomically accumulated
  float* g_ind_integral;      // [n][3]
  float* g_normal;            // [n][3] or null: gradient of the shading normal (CESR after iteration 1000 renders with
                              // normal_net's normals, training/train_cesr.py:508; the PBR stage passes a detached normal)
};

template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* sh /* [NV][4] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = warp_sum(v[i]);
    if (lane == 0) sh[i * 4 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = (sh[i * 4] + sh[i * 4 + 1]) + (sh[i * 4 + 2] + sh[i * 4 + 3]);
  __syncthreads();
}

__global__ void __launch_bounds__(128) sg_render_fwd_kernel(SgParams p) {
  __shared__ float sh[16 * 4];
  const int i = blockIdx.x, tid = threadIdx.x;
  const V3<float> nrm = {p.normal[3 * i], p.normal[3 * i + 1], p.normal[3 * i + 2]};
  const V3<float> view = {p.view[3 * i], p.view[3 * i + 1], p.view[3 * i + 2]};
  const float rough = p.rough[i];
  const float alb[3] = {p.albedo[3 * i], p.albedo[3 * i + 1], p.albedo[3 * i + 2]};
  const float sr = p.spec_refl[0];
  const SpecPoint<float> sp = spec_point<float>(nrm, view, rough);
  const float F = fresnel<float>(sr, sp.v_dot_h);
  const float bvd = p.bv_dir[i], bvi = p.bv_ind[i];
  // v[0..2] spec, v[3..5] diff, v[6..8] sum(lv*mu), v[9..11] sum(mu), v[12..14] indirect spec
  float v[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) v[k] = 0.f;
  for (int m = tid; m < p.M; m += blockDim.x) {
    const LightSG<float> l = decode_light<float>(p.lgt + 7 * m);
    const float lv = p.light_vis[(size_t)i * p.M + m];
    const float Ks = spec_lobe_kernel<float>(nrm, sp, l.lobe, l.lam);
    const float Kd = diffuse_lobe_kernel<float>(nrm, l.lobe, l.lam);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[c] += l.mu[c] * bvd * F * Ks;
      const float fm = p.lin_diff ? l.mu[c] * lv : l.mu[c] * lv * (alb[c] / kPi);
      v[3 + c] += fm * Kd;
      v[6 + c] += lv * l.mu[c];
      v[9 + c] += l.mu[c];
    }
  }
  for (int m = tid; m < p.Mi; m += blockDim.x) {
    const LightSG<float> l = decode_light<float>(p.ind_lgt + ((size_t)i * p.Mi + m) * 7);
    const float Ks = spec_lobe_kernel<float>(nrm, sp, l.lobe, l.lam);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[12 + c] += l.mu[c] * bvi * F * Ks;
  }
  block_sum<15>(v, sh);
  if (tid < 3) {
    const int c = tid;
    const float spec = fmaxf(v[c], 0.f), diff = fmaxf(v[3 + c], 0.f);
    p.pre[9 * i + c] = v[c];
    p.pre[9 * i + 3 + c] = v[3 + c];
    p.pre[9 * i + 6 + c] = v[12 + c];
    p.sg_spec[3 * i + c] = spec;
    p.sg_diff[3 * i + c] = diff;
    p.sg_rgb[3 * i + c] = spec + diff;
    p.vis_shadow[3 * i + c] = v[6 + c] / fmaxf(v[9 + c], 1e-4f);
    const float ispec = p.Mi > 0 ? fmaxf(v[12 + c], 0.f) : 0.f;
    float idiff = 0.f;
    if (p.Mi > 0) idiff = p.lin_diff ? p.ind_integral[3 * i + c] : p.ind_integral[3 * i + c] * (alb[c] / kPi);
    p.ind_spec[3 * i + c] = ispec;
    p.ind_diff[3 * i + c] = idiff;
    p.ind_rgb[3 * i + c] = ispec + idiff;
  }
}

// dual layout, specular: 0-6 raw SG, 7 rough, 8 spec_refl, 9 brdf_vis;  diffuse: 0-6 raw SG, 7 light_vis
// one specular lobe of the backward: tangents of sum_c g[c] * mu_c * bv * F * Ks with respect to the raw SG (-> g_raw) and
// to the per-point inputs (-> acc: 0 rough, 1 spec_refl, bv_slot brdf_vis, 7-9 normal)
template <bool NG, typename DS>
__device__ __forceinline__ void spec_lobe_bwd(const V3<DS>& nS, const SpecPoint<DS>& sp, const DS& F, const float* raw,
                                              float bv, const float (&g)[3], float* g_raw, int bv_slot, float* acc) {
  DS r[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) r[k] = DS::seed(raw[k], k);
  const LightSG<DS> l = decode_light<DS>(r);
  const DS Ks = spec_lobe_kernel<DS>(nS, sp, l.lobe, l.lam);
  const DS common = DS::seed(bv, 9) * F * Ks;
  DS tot(0.f);
#pragma unroll
  for (int c = 0; c < 3; ++c) tot = tot + (l.mu[c] * common) * g[c];
#pragma unroll
  for (int k = 0; k < 7; ++k) g_raw[k] = tot.d[k];
  acc[0] += tot.d[7];
  acc[1] += tot.d[8];
  acc[bv_slot] += tot.d[9];
  if constexpr (NG) {
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[7 + k] += tot.d[10 + k];
  }
}

template <typename T, bool NG, int BASE>
__device__ __forceinline__ V3<T> seed_normal(const V3<float>& n) {
  if constexpr (NG) return {T::seed(n.x, BASE), T::seed(n.y, BASE + 1), T::seed(n.z, BASE + 2)};
  else return lift3<T>(n);
}

// NG: also differentiate with respect to the shading normal (three more dual components per lobe evaluation)
template <bool NG>
__global__ void __launch_bounds__(128) sg_render_bwd_kernel(SgParams p) {
  __shared__ float sh[10 * 4];
  const int i = blockIdx.x, tid = threadIdx.x;
  const V3<float> nrm = {p.normal[3 * i], p.normal[3 * i + 1], p.normal[3 * i + 2]};
  const V3<float> view = {p.view[3 * i], p.view[3 * i + 1], p.view[3 * i + 2]};
  const float rough = p.rough[i];
  const float alb[3] = {p.albedo[3 * i], p.albedo[3 * i + 1], p.albedo[3 * i + 2]};
  const float sr = p.spec_refl[0];
  auto ld = [&](const float* g, int c) { return g ? g[3 * i + c] : 0.f; };
  float gs[3], gd[3], gis[3], gid[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // torch.clamp(x, min=0) passes the gradient where x >= 0
    gs[c] = p.pre[9 * i + c] >= 0.f ? ld(p.g_sg_rgb, c) + ld(p.g_sg_spec, c) : 0.f;
    gd[c] = p.pre[9 * i + 3 + c] >= 0.f ? ld(p.g_sg_rgb, c) + ld(p.g_sg_diff, c) : 0.f;
    gis[c] = (p.Mi > 0 && p.pre[9 * i + 6 + c] >= 0.f) ? ld(p.g_ind_rgb, c) + ld(p.g_ind_spec, c) : 0.f;
    gid[c] = p.Mi > 0 ? ld(p.g_ind_rgb, c) + ld(p.g_ind_diff, c) : 0.f;
  }
  {
    // rows without any upstream gradient (e.g. masked-out rays of the static-shape mode) contribute nothing;
    // every output was zero-initialised by the caller
    bool any = false;
#pragma unroll
    for (int c = 0; c < 3; ++c) any |= (gs[c] != 0.f) | (gd[c] != 0.f) | (gis[c] != 0.f) | (gid[c] != 0.f);
    if (!any) {
      if constexpr (NG) {
        if (tid < 3) p.g_normal[3 * i + tid] = 0.f;
      }
      return;
    }
  }
  // dual slots, specular: 0-6 raw SG, 7 rough, 8 spec_refl, 9 brdf_vis, (10-12 normal);  diffuse: 0-6 raw SG,
  // 7 light_vis, (8-10 normal)
  constexpr int NA = NG ? 10 : 7;
  typedef Dual<NG ? 13 : 10> DS;
  typedef Dual<NG ? 11 : 8> DD;
  const V3<DS> nS = seed_normal<DS, NG, 10>(nrm);
  const V3<DD> nD = seed_normal<DD, NG, 8>(nrm);
  const V3<DS> vS = lift3<DS>(view);
  const SpecPoint<DS> sp = spec_point<DS>(nS, vS, DS::seed(rough, 7));
  const DS F = fresnel<DS>(DS::seed(sr, 8), sp.v_dot_h);
  // per-point accumulators: 0 rough, 1 spec_refl, 2 bv_dir, 3 bv_ind, 4-6 albedo, (7-9 normal)
  float acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.f;


  const bool any_dir = (gs[0] != 0.f) | (gs[1] != 0.f) | (gs[2] != 0.f) | (gd[0] != 0.f) | (gd[1] != 0.f) | (gd[2] != 0.f);
  for (int m = tid; m < p.M; m += blockDim.x) {
    float g_raw[7] = {0, 0, 0, 0, 0, 0, 0};
    float g_lv = 0.f;
    if (any_dir) {
      spec_lobe_bwd<NG, DS>(nS, sp, F, p.lgt + 7 * m, p.bv_dir[i], gs, g_raw, 2, acc);
      // diffuse
      DD r[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) r[k] = DD::seed(p.lgt[7 * m + k], k);
      const LightSG<DD> l = decode_light<DD>(r);
      const DD lv = DD::seed(p.light_vis[(size_t)i * p.M + m], 7);
      const DD Kd = diffuse_lobe_kernel<DD>(nD, l.lobe, l.lam);
      DD tot(0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = p.lin_diff ? 1.f : alb[c] / kPi;
        const DD base = l.mu[c] * lv * Kd;
        tot = tot + base * (a * gd[c]);
        if (!p.lin_diff) acc[4 + c] += base.v * gd[c] / kPi;
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) g_raw[k] += tot.d[k];
      g_lv = tot.d[7];
      if constexpr (NG) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[7 + k] += tot.d[8 + k];
      }
#pragma unroll
      for (int k = 0; k < 7; ++k)
        if (g_raw[k] != 0.f) atomicAdd(p.g_lgt + 7 * m + k, g_raw[k]);
    }
    p.g_light_vis[(size_t)i * p.M + m] = g_lv;
  }
  for (int m = tid; m < p.Mi; m += blockDim.x) {
    float g_raw[7] = {0, 0, 0, 0, 0, 0, 0};
    spec_lobe_bwd<NG, DS>(nS, sp, F, p.ind_lgt + ((size_t)i * p.Mi + m) * 7, p.bv_ind[i], gis, g_raw, 3, acc);
#pragma unroll
    for (int k = 0; k < 7; ++k) p.g_ind_lgt[((size_t)i * p.Mi + m) * 7 + k] = g_raw[k];
  }
  block_sum<NA>(acc, sh);
  if (tid == 0) {
    if constexpr (NG) {
#pragma unroll
      for (int k = 0; k < 3; ++k) p.g_normal[3 * i + k] = acc[7 + k];
    }
    p.g_rough[i] = acc[0];
    if (acc[1] != 0.f) atomicAdd(p.g_spec_refl, acc[1]);
    p.g_bv_dir[i] = acc[2];
    p.g_bv_ind[i] = acc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float ga = acc[4 + c];
      float gi = 0.f;
      if (p.Mi > 0) {
        if (p.lin_diff) {
          gi = gid[c];
        } else {
          gi = gid[c] * (alb[c] / kPi);
          ga += gid[c] * p.ind_integral[3 * i + c] / kPi;
        }
      }
      p.g_albedo[3 * i + c] = ga;
      p.g_ind_integral[3 * i + c] = gi;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Hit compaction of the fixed-capacity batch (the device-side counterpart of the reference's boolean indexing,
// implicit_differentiable_renderer.py:341-347): stable partition of the N rays into hits first, misses after.
//   pos[i]   = slot of ray i,  order[slot] = ray,  n_act = number of hits,  valid[slot] = slot < n_act,
//   pts[slot] = hit ? cam + dist * dir : 0,  view[slot] = -dir of that ray.
// One CTA (N is a training batch: ~1e3 rays); block-wide exclusive scan in shared memory.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) compact_hits_kernel(int N, const unsigned char* __restrict__ hit,
                                                               const float* __restrict__ points,
                                                               const float* __restrict__ dirs, long long* __restrict__ pos,
                                                               long long* __restrict__ order, int* __restrict__ n_act,
                                                               unsigned char* __restrict__ valid, float* __restrict__ pts,
                                                               float* __restrict__ view) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  // pass 1: hit ranks (chunks of 1024 rays, running base)
  for (int c0 = 0; c0 < N; c0 += 1024) {
    const int i = c0 + tid;
    const int h = (i < N && hit[i]) ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, h);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int base = s_base;
    const int rank = base + before + __popc(b & ((1u << lane) - 1u));
    if (i < N && h) pos[i] = rank;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base = base + tot;
    }
    __syncthreads();
  }
  const int nh = s_base;
  if (tid == 0) n_act[0] = nh;
  // pass 2: misses go after the hits, in ray order: slot = nh + (i - hits before i)
  for (int i = tid; i < N; i += 1024) valid[i] = i < nh ? 1 : 0;
  __syncthreads();
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < N; c0 += 1024) {
    const int i = c0 + tid;
    const int m = (i < N && !hit[i]) ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, m);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int base = s_base;
    if (i < N && m) pos[i] = nh + base + before + __popc(b & ((1u << lane) - 1u));
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base = base + tot;
    }
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  for (int i = tid; i < N; i += 1024) {
    const long long slot = pos[i];
    order[slot] = i;
    const bool h = hit[i] != 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      pts[3 * slot + c] = h ? points[3 * i + c] : 0.f;
      view[3 * slot + c] = -dirs[3 * i + c];
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------------------------------
// SparseAE glue of the BRDF auto-encoder (model/sg_envmap_material.py:74-94, 214-232), one launch each way:
//  latent_pair: lc = sigmoid(z) (rows [0, n)), lc_r = lc + 0.01 noise (rows [n, 2n)) -- the decoder's doubled batch;
//  brdf_head:   decoder outputs y (rows [0, n)) / y_r (rows [n, 2n)) [2n][5] -> albedo = sigmoid(y[:3]),
//               roughness = 0.9 sigmoid(y3) + 0.09, metallic = 0.99 sigmoid(y4) + 0.01 and the random_xi twins
//               (xi_metallic is the bare sigmoid, sg_envmap_material.py:231).
// ------------------------------------------------------------------------------------------------------------------
__global__ void latent_pair_fwd_kernel(int total, int n32, const float* __restrict__ z, const float* __restrict__ noise,
                                       float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float s = sigmoidf_(z[i]);
  out[i] = s;
  out[n32 + i] = s + noise[i] * 0.01f;
}
__global__ void latent_pair_bwd_kernel(int total, int n32, const float* __restrict__ z, const float* __restrict__ g,
                                       float* __restrict__ g_z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float s = sigmoidf_(z[i]);
  g_z[i] = (g[i] + g[n32 + i]) * s * (1.f - s);
}
__global__ void brdf_head_fwd_kernel(int n, const float* __restrict__ y2, float* __restrict__ albedo,
                                     float* __restrict__ rough, float* __restrict__ metal, float* __restrict__ xi_albedo,
                                     float* __restrict__ xi_rough, float* __restrict__ xi_metal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* y = y2 + (size_t)i * 5;
  const float* r = y2 + (size_t)(n + i) * 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    albedo[3 * i + c] = sigmoidf_(y[c]);
    xi_albedo[3 * i + c] = sigmoidf_(r[c]);
  }
  rough[i] = sigmoidf_(y[3]) * 0.9f + 0.09f;
  metal[i] = sigmoidf_(y[4]) * 0.99f + 0.01f;
  xi_rough[i] = sigmoidf_(r[3]) * 0.9f + 0.09f;
  xi_metal[i] = sigmoidf_(r[4]);
}
// any of the six upstream gradients may be null (treated as zero)
__global__ void brdf_head_bwd_kernel(int n, const float* __restrict__ y2, const float* __restrict__ g_albedo,
                                     const float* __restrict__ g_rough, const float* __restrict__ g_metal,
                                     const float* __restrict__ g_xi_albedo, const float* __restrict__ g_xi_rough,
                                     const float* __restrict__ g_xi_metal, float* __restrict__ g_y2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* y = y2 + (size_t)i * 5;
  const float* r = y2 + (size_t)(n + i) * 5;
  float* gy = g_y2 + (size_t)i * 5;
  float* gr = g_y2 + (size_t)(n + i) * 5;
  auto ds = [](float x) { const float s = sigmoidf_(x); return s * (1.f - s); };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gy[c] = g_albedo ? g_albedo[3 * i + c] * ds(y[c]) : 0.f;
    gr[c] = g_xi_albedo ? g_xi_albedo[3 * i + c] * ds(r[c]) : 0.f;
  }
  gy[3] = g_rough ? g_rough[i] * 0.9f * ds(y[3]) : 0.f;
  gy[4] = g_metal ? g_metal[i] * 0.99f * ds(y[4]) : 0.f;
  gr[3] = g_xi_rough ? g_xi_rough[i] * 0.9f * ds(r[3]) : 0.f;
  gr[4] = g_xi_metal ? g_xi_metal[i] * ds(r[4]) : 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
// IndirctIllumNetwork lobe decoding (model/implicit_differentiable_renderer.py:207-219): per (point, lobe) the six raw
// network outputs -> [unit axis (theta = 2 pi sigmoid, phi = pi sigmoid), lambda = 30 sigmoid + 0.1, mu = relu] (7 values).
// One thread per (point, lobe); the backward recomputes the forward from the raw outputs.
// ------------------------------------------------------------------------------------------------------------------
__global__ void decode_lobes_fwd_kernel(int total, const float* __restrict__ raw, float* __restrict__ sgs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float* o = raw + (size_t)idx * 6;
  const float theta = sigmoidf_(o[0]) * 6.283185307179586f, phi = sigmoidf_(o[1]) * 3.141592653589793f;
  float st, ct, sp, cp;
  sincosf(theta, &st, &ct);
  sincosf(phi, &sp, &cp);
  float* d = sgs + (size_t)idx * 7;
  d[0] = ct * sp; d[1] = st * sp; d[2] = cp;
  d[3] = sigmoidf_(o[2]) * 30.f + 0.1f;
  d[4] = fmaxf(o[3], 0.f); d[5] = fmaxf(o[4], 0.f); d[6] = fmaxf(o[5], 0.f);
}

__global__ void decode_lobes_bwd_kernel(int total, const float* __restrict__ raw, const float* __restrict__ g_sgs,
                                        float* __restrict__ g_raw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float* o = raw + (size_t)idx * 6;
  const float* g = g_sgs + (size_t)idx * 7;
  const float a0 = sigmoidf_(o[0]), a1 = sigmoidf_(o[1]), s2 = sigmoidf_(o[2]);
  float st, ct, sp, cp;
  sincosf(a0 * 6.283185307179586f, &st, &ct);
  sincosf(a1 * 3.141592653589793f, &sp, &cp);
  const float g_theta = g[0] * (-st * sp) + g[1] * (ct * sp);
  const float g_phi = g[0] * (ct * cp) + g[1] * (st * cp) - g[2] * sp;
  float* d = g_raw + (size_t)idx * 6;
  d[0] = g_theta * 6.283185307179586f * a0 * (1.f - a0);
  d[1] = g_phi * 3.141592653589793f * a1 * (1.f - a1);
  d[2] = g[3] * 30.f * s2 * (1.f - s2);
  d[3] = o[3] > 0.f ? g[4] : 0.f;
  d[4] = o[4] > 0.f ? g[5] : 0.f;
  d[5] = o[5] > 0.f ? g[6] : 0.f;
}

}  // namespace robir

using namespace robir;

extern "C" {

// Argument block mirrors SgParams field for field (plain pointers and ints; see include/robir_b200.h).
int robir_sg_render_fwd(const SgParams* p, void* stream) {
  if (p->n == 0) return 0;
  RB_REQUIRE(p->M >= 0 && p->Mi >= 0, "sg_render_fwd: bad lobe counts");
  sg_render_fwd_kernel<<<p->n, 128, 0, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_sg_render_bwd(const SgParams* p, void* stream) {
  if (p->n == 0) return 0;
  if (p->g_normal != nullptr) sg_render_bwd_kernel<true><<<p->n, 128, 0, (cudaStream_t)stream>>>(*p);
  else sg_render_bwd_kernel<false><<<p->n, 128, 0, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// raw [n * lobes][6] -> sgs [n * lobes][7]  (and its backward); total = n * lobes
int robir_decode_lobes_fwd(int total, const float* raw, float* sgs, void* stream) {
  if (total == 0) return 0;
  decode_lobes_fwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(total, raw, sgs);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_decode_lobes_bwd(int total, const float* raw, const float* g_sgs, float* g_raw, void* stream) {
  if (total == 0) return 0;
  decode_lobes_bwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(total, raw, g_sgs, g_raw);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Stable hit-first partition of one ray batch (see compact_hits_kernel); pos / order are int64 [N].
int robir_compact_hits(int N, const unsigned char* hit, const float* points, const float* dirs, long long* pos,
                       long long* order, int* n_act, unsigned char* valid, float* pts, float* view, void* stream) {
  if (N == 0) return 0;
  compact_hits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(N, hit, points, dirs, pos, order, n_act, valid, pts, view);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// SparseAE glue of the BRDF auto-encoder (see latent_pair_* / brdf_head_* above).  z / noise [n][32]; out, g [2n][32].
int robir_latent_pair_fwd(int n, const float* z, const float* noise, float* out, void* stream) {
  if (n == 0) return 0;
  latent_pair_fwd_kernel<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n * 32, n * 32, z, noise, out);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int robir_latent_pair_bwd(int n, const float* z, const float* g, float* g_z, void* stream) {
  if (n == 0) return 0;
  latent_pair_bwd_kernel<<<(n * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n * 32, n * 32, z, g, g_z);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int robir_brdf_head_fwd(int n, const float* y2, float* albedo, float* rough, float* metal, float* xi_albedo,
                        float* xi_rough, float* xi_metal, void* stream) {
  if (n == 0) return 0;
  brdf_head_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, y2, albedo, rough, metal, xi_albedo, xi_rough,
                                                                          xi_metal);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int robir_brdf_head_bwd(int n, const float* y2, const float* g_albedo, const float* g_rough, const float* g_metal,
                        const float* g_xi_albedo, const float* g_xi_rough, const float* g_xi_metal, float* g_y2,
                        void* stream) {
  if (n == 0) return 0;
  brdf_head_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, y2, g_albedo, g_rough, g_metal, g_xi_albedo,
                                                                          g_xi_rough, g_xi_metal, g_y2);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
