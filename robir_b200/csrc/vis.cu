// Visibility path of the SG renderer: sample directions, pair compaction, the fused
// visibility-MLP evaluation over (point, direction) pairs and the weighted per-lobe reductions, forward and backward.
// Reference behaviour: model/sg_render.py:111-195 (get_diffuse_visibility), :198-301 (get_specular_visibility),
// model/implicit_differentiable_renderer.py:225-258 (VisNetwork).  This file holds the fp32 FFMA engine version
// (exact-fp32 arithmetic, the on-device cross-check for the tcgen05 version in vis_tc.cu).
#include "mlp_engine.cuh"
#include "sg_math.h"

namespace robir {

// ------------------------------------------------------------------------------------------------------------------
// weight packing: W [N][K] row-major (torch Linear) -> Wt [Kpad][Npad] with column/row windows, zero padded
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_transpose_kernel(const float* __restrict__ W, int N, int K, int k_begin, int k_count,
                                      float* __restrict__ Wt, int Kpad, int Npad, float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Kpad * Npad) return;
  const int k = idx / Npad, n = idx % Npad;
  float v = 0.f;
  if (k < k_count && n < N) v = W[(size_t)n * K + k_begin + k] * scale;
  Wt[idx] = v;
}
// row-major copy of a column window, zero padded: out[n][kpad]
__global__ void pack_window_kernel(const float* __restrict__ W, int N, int K, int k_begin, int k_count,
                                   float* __restrict__ out, int Npad, int Kpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Kpad * Npad) return;
  const int n = idx / Kpad, k = idx % Kpad;
  float v = 0.f;
  if (k < k_count && n < N) v = W[(size_t)n * K + k_begin + k];
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------------------------------
// layer-0 tables: tab[row][256] = Wt[64][256]^T . PE10(x[row]) (+ bias)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pe10_to_tile(float* Xs, int RP, int r, float x, float y, float z, bool valid) {
  // k-major tile rows 0..62 = [x, sin(2^f x), cos(2^f x)]_f, row 63 = 0    (model/embedder.py:24-38)
  const float v[3] = {valid ? x : 0.f, valid ? y : 0.f, valid ? z : 0.f};
#pragma unroll
  for (int i = 0; i < 3; ++i) Xs[i * RP + r] = v[i];
  float f = 1.f;
  for (int l = 0; l < 10; ++l) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float a = v[i] * f;
      Xs[(3 + 6 * l + i) * RP + r] = valid ? sinf(a) : 0.f;
      Xs[(3 + 6 * l + 3 + i) * RP + r] = valid ? cosf(a) : 0.f;
    }
    f *= 2.f;
  }
  Xs[63 * RP + r] = 0.f;
}

__global__ void __launch_bounds__(256) pe_linear_kernel(const float* __restrict__ x, int n, const float* __restrict__ Wt,
                                                        const float* __restrict__ bias, float* __restrict__ tab) {
  constexpr int R = 64, RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                 // [64][RP]
  float* Wbuf = smem + 64 * RP;     // 32 KB
  const int row0 = blockIdx.x * R;
  if (threadIdx.x < R) {
    const int r = row0 + threadIdx.x;
    const bool ok = r < n;
    pe10_to_tile(Xs, RP, threadIdx.x, ok ? x[3 * r] : 0.f, ok ? x[3 * r + 1] : 0.f, ok ? x[3 * r + 2] : 0.f, ok);
  }
  __syncthreads();
  float acc[TR][8];
  zero_acc<R>(acc);
  tile_gemm_pass<R>(Xs, 64, Wt, 256, 0, Wbuf, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 b0 = make_float4(0, 0, 0, 0), b1 = b0;
  if (bias) {
    b0 = *reinterpret_cast<const float4*>(bias + lane * 4);
    b1 = *reinterpret_cast<const float4*>(bias + 128 + lane * 4);
  }
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    const int row = row0 + warp * TR + r;
    if (row < n) {
      float* dst = tab + (size_t)row * 256;
      *reinterpret_cast<float4*>(dst + lane * 4) =
          make_float4(acc[r][0] + b0.x, acc[r][1] + b0.y, acc[r][2] + b0.z, acc[r][3] + b0.w);
      *reinterpret_cast<float4*>(dst + 128 + lane * 4) =
          make_float4(acc[r][4] + b1.x, acc[r][5] + b1.y, acc[r][6] + b1.z, acc[r][7] + b1.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// sample directions.  K axes, S samples each.  Diffuse: axis_f = axis_w = norm_axis(lobe) (computed here from the
// once-normalised lobe), sharp = clamp(lam, 1e-4), lam_w = lam.  Specular: axis_f = reflection dir, axis_w = warped
// lobe, sharp = lam_w = clip(warp lambda, 0.1, 50) -- all prepared by the caller.
// ------------------------------------------------------------------------------------------------------------------
__global__ void sample_dirs_fwd_kernel(int K, int S, const float* __restrict__ axis_f, const float* __restrict__ axis_w,
                                       const float* __restrict__ sharp, const float* __restrict__ lam_w,
                                       const float* __restrict__ sg_range, const float* __restrict__ u_theta,
                                       const float* __restrict__ u_phi, int renorm_axis, float* __restrict__ dirs,
                                       float* __restrict__ w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * S) return;
  const int k = idx / S;
  V3<float> af = {axis_f[3 * k], axis_f[3 * k + 1], axis_f[3 * k + 2]};
  V3<float> aw = {axis_w[3 * k], axis_w[3 * k + 1], axis_w[3 * k + 2]};
  if (renorm_axis) { af = norm_axis(af); aw = af; }
  V3<float> d;
  float ww;
  sample_dir<float>(af, aw, sharp[k], lam_w[k], sg_range[0], u_theta[idx], u_phi[idx], &d, &ww);
  dirs[3 * idx] = d.x; dirs[3 * idx + 1] = d.y; dirs[3 * idx + 2] = d.z;
  w[idx] = ww;
}

// backward: g_dirs [K*S,3], g_w [K*S] -> g_axis_f [K,3], g_axis_w [K,3], g_sharp [K], g_lam_w [K], g_sg_range [1]
__global__ void sample_dirs_bwd_kernel(int K, int S, const float* __restrict__ axis_f, const float* __restrict__ axis_w,
                                       const float* __restrict__ sharp, const float* __restrict__ lam_w,
                                       const float* __restrict__ sg_range, const float* __restrict__ u_theta,
                                       const float* __restrict__ u_phi, int renorm_axis,
                                       const float* __restrict__ g_dirs, const float* __restrict__ g_w,
                                       float* __restrict__ g_axis_f, float* __restrict__ g_axis_w,
                                       float* __restrict__ g_sharp, float* __restrict__ g_lam_w,
                                       float* __restrict__ g_sg_range) {
  // one warp per axis; lanes stride over samples; inputs: 0-2 axis_f, 3-5 axis_w, 6 sharp, 7 lam_w, 8 sg_range
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  typedef Dual<9> D;
  float g[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = 0.f;
  for (int s = lane; s < S; s += 32) {
    const int idx = k * S + s;
    V3<D> af = {D::seed(axis_f[3 * k], 0), D::seed(axis_f[3 * k + 1], 1), D::seed(axis_f[3 * k + 2], 2)};
    V3<D> aw;
    if (renorm_axis) {
      af = norm_axis(af);
      aw = af;
    } else {
      aw = {D::seed(axis_w[3 * k], 3), D::seed(axis_w[3 * k + 1], 4), D::seed(axis_w[3 * k + 2], 5)};
    }
    V3<D> d;
    D ww;
    sample_dir<D>(af, aw, D::seed(sharp[k], 6), D::seed(lam_w[k], 7), D::seed(sg_range[0], 8), u_theta[idx],
                  u_phi[idx], &d, &ww);
    const float gx = g_dirs[3 * idx], gy = g_dirs[3 * idx + 1], gz = g_dirs[3 * idx + 2], gw = g_w[idx];
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] += gx * d.x.d[i] + gy * d.y.d[i] + gz * d.z.d[i] + gw * ww.d[i];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = warp_sum(g[i]);
  if (lane == 0) {
    for (int i = 0; i < 3; ++i) g_axis_f[3 * k + i] = g[i];
    if (g_axis_w) for (int i = 0; i < 3; ++i) g_axis_w[3 * k + i] = g[3 + i];
    g_sharp[k] = g[6];
    g_lam_w[k] = g[7];
    atomicAdd(g_sg_range, g[8]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// diffuse pair compaction: live(i, j) = n_i . dir_j > 1e-6   (sg_render.py:155).  S <= 32 samples per lobe.
// ------------------------------------------------------------------------------------------------------------------
__global__ void diffuse_count_kernel(int n, int M, int S, const float* __restrict__ normals,
                                     const float* __restrict__ dirs, uint32_t* __restrict__ bits,
                                     int* __restrict__ lobe_off /*[n][M+1]*/) {
  // one CTA per point; warp per lobe (ballot over its S samples)
  extern __shared__ int s_cnt[];  // [M]
  const int i = blockIdx.x;
  const float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int m = warp; m < M; m += nwarp) {
    bool live = false;
    if (lane < S) {
      const float* d = dirs + 3 * (m * S + lane);
      // torch.sum(normals * dir, -1): three products, sequential adds, no contraction
      const float dp = __fadd_rn(__fadd_rn(__fmul_rn(nx, d[0]), __fmul_rn(ny, d[1])), __fmul_rn(nz, d[2]));
      live = dp > kTiny;
    }
    const unsigned b = __ballot_sync(0xffffffffu, live);
    if (lane == 0) {
      bits[(size_t)i * M + m] = b;
      s_cnt[m] = __popc(b);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int m = 0; m < M; ++m) {
      lobe_off[(size_t)i * (M + 1) + m] = run;
      run += s_cnt[m];
    }
    lobe_off[(size_t)i * (M + 1) + M] = run;
  }
}

// exclusive scan of 64-padded per-point counts; writes start[i] (row offset) and counters: n_tiles, n_pairs
// pad = tile: every point is padded to a tile multiple; pad = 1: points are packed back to back and only the last
// tile is padded (rowB = -1 on its tail), which saves ~half a tile of dead rows per point
__global__ void diffuse_scan_kernel(int n, int M, int pad, int tile, const int* __restrict__ lobe_off,
                                    int* __restrict__ start, int* __restrict__ n_tiles, long long* __restrict__ n_pairs,
                                    int* __restrict__ rowA, int* __restrict__ rowB) {
  __shared__ int s_total;
  __shared__ int s_part[1024];
  __shared__ long long s_pairs[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (n + T - 1) / T;
  int sum = 0;
  long long pairs = 0;
  for (int q = 0; q < per; ++q) {
    const int i = t * per + q;
    if (i < n) {
      const int c = lobe_off[(size_t)i * (M + 1) + M];
      sum += ((c + pad - 1) / pad) * pad;
      pairs += c;
    }
  }
  s_part[t] = sum;
  s_pairs[t] = pairs;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    long long p = 0;
    for (int q = 0; q < T; ++q) {
      const int v = s_part[q];
      s_part[q] = run;
      run += v;
      p += s_pairs[q];
    }
    *n_tiles = (run + tile - 1) / tile;
    *n_pairs += p;
    s_total = run;
  }
  __syncthreads();
  for (int q = s_total + t; q < ((s_total + tile - 1) / tile) * tile; q += T) {   // tail of the last tile
    rowA[q] = 0;
    rowB[q] = -1;
  }
  int run = s_part[t];
  for (int q = 0; q < per; ++q) {
    const int i = t * per + q;
    if (i < n) {
      start[i] = run;
      run += ((lobe_off[(size_t)i * (M + 1) + M] + pad - 1) / pad) * pad;
    }
  }
}

__global__ void diffuse_fill_kernel(int n, int M, int S, int pad, const uint32_t* __restrict__ bits,
                                    const int* __restrict__ lobe_off, const int* __restrict__ start,
                                    int* __restrict__ rowA, int* __restrict__ rowB) {
  const int i = blockIdx.x;
  const int base = start[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int m = warp; m < M; m += nwarp) {
    const unsigned b = bits[(size_t)i * M + m];
    if ((b >> lane) & 1u) {
      const int q = base + lobe_off[(size_t)i * (M + 1) + m] + __popc(b & ((1u << lane) - 1u));
      rowA[q] = i;
      rowB[q] = m * S + lane;
    }
  }
  const int cnt = lobe_off[(size_t)i * (M + 1) + M];
  const int padded = ((cnt + pad - 1) / pad) * pad;
  for (int q = cnt + threadIdx.x; q < padded; q += blockDim.x) {
    rowA[base + q] = i;
    rowB[base + q] = -1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fused visibility MLP, forward.  One 64-row tile per loop iteration:
//   h1 = relu(tabA[a] + tabB[b]); h2..h4 = relu(W h + b); v = sigmoid((w1-w0).h4 + (b1-b0))
// ------------------------------------------------------------------------------------------------------------------
struct VisFwdParams {
  const float* tabA; const float* tabB;
  const int* rowA; const int* rowB;
  const int* n_tiles;
  const float* Wt[3];   // packed transposed [256][256]
  const float* bias[3];
  const float* wd;      // [256] = W4[1]-W4[0]
  const float* bd;      // [1]
  float* vis;           // [rows]
  uint32_t* mask;       // [rows][4][8] or null
};

__global__ void __launch_bounds__(256, 2) vis_mlp_fwd_kernel(VisFwdParams p) {
  constexpr int R = 64, RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                     // [256][RP]
  float* Wbuf = Xs + 256 * RP;          // 32 KB
  float* red = Wbuf + kWbufFloats;      // [4][64]
  __shared__ int s_a[R], s_b[R];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = *p.n_tiles;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int q0 = tile * R;
    __syncthreads();
    if (tid < R) {
      s_a[tid] = p.rowA[q0 + tid];
      s_b[tid] = p.rowB[q0 + tid];
    }
    __syncthreads();
    float acc[TR][8];
    // ---- stage 0: gather-add the layer-0 halves
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int row = warp * TR + r;
      const int a = s_a[row], b = s_b[row];
      if (b >= 0) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.tabA + (size_t)a * 256 + lane * 4));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.tabA + (size_t)a * 256 + 128 + lane * 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.tabB + (size_t)b * 256 + lane * 4));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.tabB + (size_t)b * 256 + 128 + lane * 4));
        acc[r][0] = a0.x + b0.x; acc[r][1] = a0.y + b0.y; acc[r][2] = a0.z + b0.z; acc[r][3] = a0.w + b0.w;
        acc[r][4] = a1.x + b1.x; acc[r][5] = a1.y + b1.y; acc[r][6] = a1.z + b1.z; acc[r][7] = a1.w + b1.w;
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
      }
    }
    for (int layer = 0; layer < 4; ++layer) {
      if (layer > 0) {
        zero_acc<R>(acc);
        tile_gemm_pass<R>(Xs, 256, p.Wt[layer - 1], 256, 0, Wbuf, acc);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias[layer - 1] + lane * 4));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias[layer - 1] + 128 + lane * 4));
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] += bb[c];
      }
      if (p.mask)
        store_relu_mask<R>(acc, [&](int row) -> uint32_t* {
          return s_b[row] >= 0 ? p.mask + ((size_t)(q0 + row) * 4 + layer) * 8 : nullptr;
        });
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaxf(acc[r][c], 0.f);
      store_acc<R>(Xs, 0, 256, acc);
      __syncthreads();
    }
    // ---- logit difference + sigmoid
    {
      const int row = tid & 63, part = tid >> 6;
      float s = 0.f;
      for (int k = part * 64; k < part * 64 + 64; ++k) s = fmaf(__ldg(p.wd + k), Xs[k * RP + row], s);
      red[part * 64 + row] = s;
    }
    __syncthreads();
    if (tid < R) {
      const float z = ((red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid])) + __ldg(p.bd);
      p.vis[q0 + tid] = s_b[tid] >= 0 ? 1.f / (1.f + expf(-z)) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fused visibility MLP, backward w.r.t. the direction input (weights and points get no gradient on this path:
// points are detached, sg_render.py:376-378; VisNetwork parameters are not optimised in the PBR stage,
// training/train_pbr.py:104-105).    g_vis[q] is dL/dv of row q.
// ------------------------------------------------------------------------------------------------------------------
struct VisBwdParams {
  const int* rowB;
  const int* n_tiles;
  const float* W[3];     // W1..W3 row-major [256][256]
  const float* W0d;      // [256][64] = W0[:, 63:126] zero padded
  const float* wd;
  const float* vis; const float* g_vis;
  const uint32_t* mask;
  const float* dirs;     // [nB][3]
  float* g_dirs;         // [nB][3], accumulated atomically
};

__global__ void __launch_bounds__(256, 2) vis_mlp_bwd_kernel(VisBwdParams p) {
  constexpr int R = 64, RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Wbuf = Xs + 256 * RP;
  __shared__ int s_b[R];
  __shared__ float s_g[R];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = *p.n_tiles;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int q0 = tile * R;
    __syncthreads();
    if (tid < R) {
      const int b = p.rowB[q0 + tid];
      s_b[tid] = b;
      const float v = p.vis[q0 + tid];
      s_g[tid] = b >= 0 ? p.g_vis[q0 + tid] * v * (1.f - v) : 0.f;
    }
    __syncthreads();
    float acc[TR][8];
    const uint32_t* mrow[TR];
    // ---- dL/dh4 (post-ReLU mask of layer 3 = mask index 3)
    {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.wd + lane * 4));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.wd + 128 + lane * 4));
      const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const int row = warp * TR + r;
        mrow[r] = s_b[row] >= 0 ? p.mask + ((size_t)(q0 + row) * 4 + 3) * 8 : nullptr;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = s_g[row] * wr[c];
      }
      apply_relu_mask<R>(acc, mrow);
      store_acc<R>(Xs, 0, 256, acc);
      __syncthreads();
    }
    for (int layer = 2; layer >= 0; --layer) {   // through W3, W2, W1
      zero_acc<R>(acc);
      tile_gemm_pass<R>(Xs, 256, p.W[layer], 256, 0, Wbuf, acc);
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const int row = warp * TR + r;
        mrow[r] = s_b[row] >= 0 ? p.mask + ((size_t)(q0 + row) * 4 + layer) * 8 : nullptr;
      }
      apply_relu_mask<R>(acc, mrow);
      store_acc<R>(Xs, 0, 256, acc);
      __syncthreads();
    }
    // ---- through the direction half of layer 0: dPE [64] (only columns 0..63 of the pass are non-zero)
    zero_acc<R>(acc);
    tile_gemm_pass<R>(Xs, 256, p.W0d, 256, 0, Wbuf, acc);   // W0d is zero padded to [256][256]
    store_acc<R>(Xs, 0, 64, acc);
    __syncthreads();
    // ---- PE jacobian: d dir_i = dPE[i] + sum_f 2^f (cos(2^f x_i) dPE[3+6f+i] - sin(2^f x_i) dPE[6+6f+i])
    if (tid < R * 3) {
      const int row = tid / 3, i = tid % 3;
      const int b = s_b[row];
      if (b >= 0) {
        const float x = __ldg(p.dirs + 3 * b + i);
        float g = Xs[i * RP + row];
        float f = 1.f;
        for (int l = 0; l < 10; ++l) {
          float sn, cs;
          sincosf(x * f, &sn, &cs);
          g += f * (cs * Xs[(3 + 6 * l + i) * RP + row] - sn * Xs[(6 + 6 * l + i) * RP + row]);
          f *= 2.f;
        }
        atomicAdd(p.g_dirs + 3 * b + i, g);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------------------------
// diffuse forward: LV[i][m] = sum_s vis * w[m,s] / (sum_s w[m,s] + 1e-6)        (sg_render.py:180-183)
__global__ void diffuse_reduce_fwd_kernel(int n, int M, int S, const uint32_t* __restrict__ bits,
                                          const int* __restrict__ lobe_off, const int* __restrict__ start,
                                          const float* __restrict__ vis, const float* __restrict__ w,
                                          float* __restrict__ light_vis /*[n][M]*/) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * M) return;
  const int i = idx / M, m = idx % M;
  const unsigned b = bits[idx];
  int q = start[i] + lobe_off[(size_t)i * (M + 1) + m];
  float num = 0.f, den = 0.f;
  for (int s = 0; s < S; ++s) {
    const float ws = w[m * S + s];
    den += ws;
    if ((b >> s) & 1u) num += vis[q++] * ws;
  }
  light_vis[idx] = num / (den + kTiny);
}

// diffuse backward: g_vis[q] and g_w[m,s] from g_LV[i][m].  One CTA per lobe m, threads over the points: the S
// per-sample sums over the points are reduced inside the CTA and written once (no atomics, fixed summation order).
__global__ void __launch_bounds__(256) diffuse_reduce_bwd_kernel(int n, int M, int S, const uint32_t* __restrict__ bits,
                                                                 const int* __restrict__ lobe_off,
                                                                 const int* __restrict__ start,
                                                                 const float* __restrict__ vis, const float* __restrict__ w,
                                                                 const float* __restrict__ light_vis,
                                                                 const float* __restrict__ g_lv, float* __restrict__ g_vis,
                                                                 float* __restrict__ g_w) {
  __shared__ float s_w[32];
  __shared__ float s_red[8][32];
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 32) s_w[tid] = tid < S ? w[m * S + tid] : 0.f;
  __syncthreads();
  float den = 0.f;
  for (int s = 0; s < S; ++s) den += s_w[s];
  const float inv = 1.f / (den + kTiny);
  float acc[32];
#pragma unroll
  for (int s = 0; s < 32; ++s) acc[s] = 0.f;
  for (int i = tid; i < n; i += blockDim.x) {
    const size_t idx = (size_t)i * M + m;
    const float g = g_lv[idx];
    const unsigned b = bits[idx];
    if (g == 0.f && b == 0u) continue;
    int q = start[i] + lobe_off[(size_t)i * (M + 1) + m];
    const float lv = light_vis[idx];
#pragma unroll
    for (int s = 0; s < 32; ++s) {
      if (s >= S) break;
      float v = 0.f;
      if ((b >> s) & 1u) {
        v = vis[q];
        g_vis[q] = g * s_w[s] * inv;
        ++q;
      }
      acc[s] += g * (v - lv) * inv;
    }
  }
#pragma unroll
  for (int s = 0; s < 32; ++s) {
    const float v = warp_sum(acc[s]);
    if (lane == 0) s_red[warp][s] = v;
  }
  __syncthreads();
  if (tid < S) {
    float t = 0.f;
    for (int wp = 0; wp < 8; ++wp) t += s_red[wp][tid];
    g_w[m * S + tid] = t;
  }
}

// specular: rows q = i*S + s (fixed layout).  fwd: BV[i] = sum_s v' w / (sum_s w + 1e-6), v' = inv ? 1 - v : v
__global__ void spec_reduce_fwd_kernel(int n, int S, int inv, int testing, const int* __restrict__ rowB,
                                       const float* __restrict__ vis, const float* __restrict__ w,
                                       float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float num = 0.f, den = 0.f;
  bool any_inf = false;
  for (int s = 0; s < S; ++s) any_inf |= isinf(w[i * S + s]);
  for (int s = 0; s < S; ++s) {
    float ws = w[i * S + s];
    if (testing && any_inf) ws = isinf(ws) ? 1.f : 0.f;   // sg_render.py:285-292
    den += ws;
    if (rowB[i * S + s] >= 0) {
      const float v = vis[i * S + s];
      num += (inv ? 1.f - v : v) * ws;
    }
  }
  out[i] = num / (den + kTiny);
}
__global__ void spec_reduce_bwd_kernel(int n, int S, int inv, const int* __restrict__ rowB,
                                       const float* __restrict__ vis, const float* __restrict__ w,
                                       const float* __restrict__ out, const float* __restrict__ g_out,
                                       float* __restrict__ g_vis, float* __restrict__ g_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float den = 0.f;
  for (int s = 0; s < S; ++s) den += w[i * S + s];
  const float invd = 1.f / (den + kTiny);
  const float g = g_out[i], o = out[i];
  for (int s = 0; s < S; ++s) {
    const int q = i * S + s;
    float v = 0.f;
    if (rowB[q] >= 0) {
      v = inv ? 1.f - vis[q] : vis[q];
      g_vis[q] = (inv ? -1.f : 1.f) * g * w[q] * invd;
    } else {
      g_vis[q] = 0.f;
    }
    g_w[q] = g * (v - o) * invd;
  }
}

// specular row list: rowB[q] = q if n_i . dir_q > 1e-6 else -1; rows beyond n*S (tile padding) = -1
// a_mod > 0: the list concatenates copies of the same a_mod points (direct + indirect BRDF-lobe queries in one launch)
__global__ void spec_rows_kernel(int n, int S, int rows_padded, int pad, int a_mod, const float* __restrict__ normals,
                                 const float* __restrict__ dirs, int* __restrict__ rowA, int* __restrict__ rowB,
                                 int* __restrict__ n_tiles, long long* __restrict__ n_pairs) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) *n_tiles = rows_padded / pad;
  if (q >= rows_padded) return;
  int a = 0, b = -1;
  if (q < n * S) {
    a = q / S;
    if (a_mod > 0) a %= a_mod;
    const float* d = dirs + 3 * q;
    const float dp = __fadd_rn(__fadd_rn(__fmul_rn(normals[3 * a], d[0]), __fmul_rn(normals[3 * a + 1], d[1])),
                               __fmul_rn(normals[3 * a + 2], d[2]));
    if (dp > kTiny) b = q;
  }
  rowA[q] = a;
  rowB[q] = b;
  const unsigned live = __ballot_sync(__activemask(), b >= 0);
  if ((threadIdx.x & 31) == 0 && live) atomicAdd((unsigned long long*)n_pairs, (unsigned long long)__popc(live));
}

// ------------------------------------------------------------------------------------------------------------------
// BRDF-lobe sampling inputs for get_specular_visibility (model/sg_render.py:198-225) from the warp of render_with_sg
// (:417-428), once for the direct and the indirect call: per point
//   vdl = clamp(n.v, 0), ref = 2 vdl n - v (reflection axis), wl = ref / (|ref| + 1e-6), wlam = (2 / r^4) / (4 vdl + 1e-6),
//   sharp = clip(wlam, 0.1, 50);   sg_range = clamp(min over the valid points of sharp, max = 1)  (batch-global, :220-222)
// Outputs are written twice (rows i and n + i) so that both calls share one launch chain.  One CTA.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) spec_prep_fwd_kernel(int n, const float* __restrict__ normal,
                                                                const float* __restrict__ view,
                                                                const float* __restrict__ rough,
                                                                const unsigned char* __restrict__ valid,
                                                                float* __restrict__ ref, float* __restrict__ wl,
                                                                float* __restrict__ sharp, float* __restrict__ sg_range,
                                                                float* __restrict__ wlam_out, int* __restrict__ argmin) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  float best = INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float nx = normal[3 * i], ny = normal[3 * i + 1], nz = normal[3 * i + 2];
    const float vx = view[3 * i], vy = view[3 * i + 1], vz = view[3 * i + 2];
    const float r = rough[i];
    const float vdl = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(nx, vx), __fmul_rn(ny, vy)), __fmul_rn(nz, vz)), 0.f);
    const float rx = 2.f * vdl * nx - vx, ry = 2.f * vdl * ny - vy, rz = 2.f * vdl * nz - vz;
    const float inv = 1.f / (sqrtf(rx * rx + ry * ry + rz * rz) + 1e-6f);
    const float wlam = (2.f / (r * r * r * r)) / (4.f * vdl + 1e-6f);
    const float sh = fminf(fmaxf(wlam, 0.1f), 50.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t o = (size_t)h * n + i;
      ref[3 * o] = rx; ref[3 * o + 1] = ry; ref[3 * o + 2] = rz;
      wl[3 * o] = rx * inv; wl[3 * o + 1] = ry * inv; wl[3 * o + 2] = rz * inv;
      sharp[o] = sh;
    }
    wlam_out[i] = wlam;
    if ((valid == nullptr || valid[i]) && (sh < best || (sh == best && i < best_i))) { best = sh; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov < best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = best_i; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_v[w] < best || (s_v[w] == best && s_i[w] < best_i)) { best = s_v[w]; best_i = s_i[w]; }
    sg_range[0] = fminf(best, 1.f);
    argmin[0] = (best <= 1.f && best_i != 0x7fffffff) ? best_i : -1;   // row that receives d / d sg_range
  }
}

// g_rough[i] = (g_sharp[i] + g_sharp[n + i] + [i == argmin] g_sg_range) * [0.1 <= wlam <= 50] * d wlam / d r
__global__ void spec_prep_bwd_kernel(int n, const float* __restrict__ rough, const float* __restrict__ wlam,
                                     const int* __restrict__ argmin, const float* __restrict__ g_sharp,
                                     const float* __restrict__ g_sg_range, float* __restrict__ g_rough) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float w = wlam[i];
  float g = 0.f;
  if (w >= 0.1f && w <= 50.f) {
    g = g_sharp[i] + g_sharp[n + i];
    if (argmin[0] == i) g += g_sg_range[0];
    g *= -4.f * w / rough[i];
  }
  g_rough[i] = g;
}

}  // namespace robir

// ======================================================================================================================
// C ABI
// ======================================================================================================================
using namespace robir;

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static constexpr int kVisSmem = (256 * (64 + 4) + kWbufFloats + 256) * 4;

extern "C" {

int robir_pack_transpose(const float* W, int N, int K, int k_begin, int k_count, float* Wt, int Kpad, int Npad,
                         float scale, void* stream) {
  RB_REQUIRE(k_begin >= 0 && k_begin + k_count <= K && k_count <= Kpad && N <= Npad, "pack_transpose: bad window");
  pack_transpose_kernel<<<cdiv((long long)Kpad * Npad, 256), 256, 0, (cudaStream_t)stream>>>(W, N, K, k_begin, k_count,
                                                                                              Wt, Kpad, Npad, scale);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_pack_window(const float* W, int N, int K, int k_begin, int k_count, float* out, int Npad, int Kpad,
                      void* stream) {
  RB_REQUIRE(k_begin >= 0 && k_begin + k_count <= K && k_count <= Kpad && N <= Npad, "pack_window: bad window");
  pack_window_kernel<<<cdiv((long long)Kpad * Npad, 256), 256, 0, (cudaStream_t)stream>>>(W, N, K, k_begin, k_count, out,
                                                                                           Npad, Kpad);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// tab[n][256] = PE10(x[n][3]) . Wt[64][256] (+ bias[256] or null)
int robir_pe_linear(const float* x, int n, const float* Wt, const float* bias, float* tab, void* stream) {
  if (n == 0) return 0;
  const int smem = (64 * 68 + kWbufFloats) * 4;
  RB_CHECK_CUDA(cudaFuncSetAttribute(pe_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pe_linear_kernel<<<cdiv(n, 64), 256, smem, (cudaStream_t)stream>>>(x, n, Wt, bias, tab);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_sample_dirs_fwd(int K, int S, const float* axis_f, const float* axis_w, const float* sharp,
                          const float* lam_w, const float* sg_range, const float* u_theta, const float* u_phi,
                          int renorm_axis, float* dirs, float* w, void* stream) {
  if (K * S == 0) return 0;
  sample_dirs_fwd_kernel<<<cdiv((long long)K * S, 256), 256, 0, (cudaStream_t)stream>>>(
      K, S, axis_f, axis_w, sharp, lam_w, sg_range, u_theta, u_phi, renorm_axis, dirs, w);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// g_sg_range must be zero-initialised by the caller; g_axis_w may be null when renorm_axis != 0
int robir_sample_dirs_bwd(int K, int S, const float* axis_f, const float* axis_w, const float* sharp,
                          const float* lam_w, const float* sg_range, const float* u_theta, const float* u_phi,
                          int renorm_axis, const float* g_dirs, const float* g_w, float* g_axis_f, float* g_axis_w,
                          float* g_sharp, float* g_lam_w, float* g_sg_range, void* stream) {
  if (K * S == 0) return 0;
  sample_dirs_bwd_kernel<<<cdiv(K, 4), 128, 0, (cudaStream_t)stream>>>(K, S, axis_f, axis_w, sharp, lam_w, sg_range,
                                                                       u_theta, u_phi, renorm_axis, g_dirs, g_w,
                                                                       g_axis_f, g_axis_w, g_sharp, g_lam_w, g_sg_range);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Builds the diffuse row list for tiles of tile_rows rows (64: FFMA engine, 128: tensor-core engine); pad_points != 0
// pads every point to a tile multiple (one point per tile), 0 packs the points back to back (tensor-core engine).  Workspace (caller-owned, int32 unless noted): bits [n*M] (u32), lobe_off [n*(M+1)], start [n],
// rowA/rowB [n * roundup(M*S, tile_rows)], counters: n_tiles [1] (int), n_pairs [1] (int64, accumulated).
int robir_diffuse_rows(int n, int M, int S, int tile_rows, int pad_points, const float* normals, const float* dirs,
                       uint32_t* bits, int* lobe_off, int* start, int* rowA, int* rowB, int* n_tiles, long long* n_pairs,
                       void* stream) {
  RB_REQUIRE(S >= 1 && S <= 32, "diffuse_rows: S must be in [1,32]");
  RB_REQUIRE(tile_rows == 64 || tile_rows == 128, "diffuse_rows: tile_rows must be 64 or 128");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    RB_CHECK_CUDA(cudaMemsetAsync(n_tiles, 0, sizeof(int), st));
    return 0;
  }
  diffuse_count_kernel<<<n, 256, M * sizeof(int), st>>>(n, M, S, normals, dirs, bits, lobe_off);
  const int pad = pad_points ? tile_rows : 1;
  diffuse_scan_kernel<<<1, 1024, 0, st>>>(n, M, pad, tile_rows, lobe_off, start, n_tiles, n_pairs, rowA, rowB);
  diffuse_fill_kernel<<<n, 256, 0, st>>>(n, M, S, pad, bits, lobe_off, start, rowA, rowB);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_spec_rows(int n, int S, int rows_padded, int tile_rows, int a_mod, const float* normals, const float* dirs,
                    int* rowA, int* rowB, int* n_tiles, long long* n_pairs, void* stream) {
  RB_REQUIRE((tile_rows == 64 || tile_rows == 128) && rows_padded % tile_rows == 0 && rows_padded >= n * S,
             "spec_rows: rows_padded must be a multiple of tile_rows (64 or 128)");
  if (rows_padded == 0) {
    RB_CHECK_CUDA(cudaMemsetAsync(n_tiles, 0, sizeof(int), (cudaStream_t)stream));
    return 0;
  }
  spec_rows_kernel<<<cdiv(rows_padded, 256), 256, 0, (cudaStream_t)stream>>>(n, S, rows_padded, tile_rows, a_mod,
                                                                             normals, dirs, rowA, rowB, n_tiles, n_pairs);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Fused visibility MLP forward over a row list (fp32 FFMA engine).  max_tiles bounds the grid; the actual tile
// count is read on the device from n_tiles.  mask may be null (no backward needed).
int robir_vis_mlp_fwd(const float* tabA, const float* tabB, const int* rowA, const int* rowB, const int* n_tiles,
                      int max_tiles, const float* Wt1, const float* Wt2, const float* Wt3, const float* b1,
                      const float* b2, const float* b3, const float* wd, const float* bd, float* vis, uint32_t* mask,
                      int sm_count, void* stream) {
  if (max_tiles == 0) return 0;
  VisFwdParams p;
  p.tabA = tabA; p.tabB = tabB; p.rowA = rowA; p.rowB = rowB; p.n_tiles = n_tiles;
  p.Wt[0] = Wt1; p.Wt[1] = Wt2; p.Wt[2] = Wt3;
  p.bias[0] = b1; p.bias[1] = b2; p.bias[2] = b3;
  p.wd = wd; p.bd = bd; p.vis = vis; p.mask = mask;
  RB_CHECK_CUDA(cudaFuncSetAttribute(vis_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVisSmem));
  const int grid = max_tiles < 2 * sm_count ? max_tiles : 2 * sm_count;
  vis_mlp_fwd_kernel<<<grid, 256, kVisSmem, (cudaStream_t)stream>>>(p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// g_dirs must be zero-initialised by the caller.
int robir_vis_mlp_bwd(const int* rowB, const int* n_tiles, int max_tiles, const float* W1, const float* W2,
                      const float* W3, const float* W0d, const float* wd, const float* vis, const float* g_vis,
                      const uint32_t* mask, const float* dirs, float* g_dirs, int sm_count, void* stream) {
  if (max_tiles == 0) return 0;
  VisBwdParams p;
  p.rowB = rowB; p.n_tiles = n_tiles;
  p.W[0] = W1; p.W[1] = W2; p.W[2] = W3;
  p.W0d = W0d; p.wd = wd; p.vis = vis; p.g_vis = g_vis; p.mask = mask; p.dirs = dirs; p.g_dirs = g_dirs;
  RB_CHECK_CUDA(cudaFuncSetAttribute(vis_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVisSmem));
  const int grid = max_tiles < 2 * sm_count ? max_tiles : 2 * sm_count;
  vis_mlp_bwd_kernel<<<grid, 256, kVisSmem, (cudaStream_t)stream>>>(p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_diffuse_reduce_fwd(int n, int M, int S, const uint32_t* bits, const int* lobe_off, const int* start,
                             const float* vis, const float* w, float* light_vis, void* stream) {
  if (n * M == 0) return 0;
  diffuse_reduce_fwd_kernel<<<cdiv((long long)n * M, 128), 128, 0, (cudaStream_t)stream>>>(n, M, S, bits, lobe_off,
                                                                                           start, vis, w, light_vis);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// g_w [M*S] is written (not accumulated); g_vis rows of dead samples are not touched (zero-initialise).
int robir_diffuse_reduce_bwd(int n, int M, int S, const uint32_t* bits, const int* lobe_off, const int* start,
                             const float* vis, const float* w, const float* light_vis, const float* g_lv, float* g_vis,
                             float* g_w, void* stream) {
  if (n * M == 0) return 0;
  diffuse_reduce_bwd_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(n, M, S, bits, lobe_off, start, vis, w, light_vis, g_lv,
                                                                 g_vis, g_w);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_spec_reduce_fwd(int n, int S, int inv, int testing, const int* rowB, const float* vis, const float* w,
                          float* out, void* stream) {
  if (n == 0) return 0;
  spec_reduce_fwd_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(n, S, inv, testing, rowB, vis, w, out);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_spec_reduce_bwd(int n, int S, int inv, const int* rowB, const float* vis, const float* w, const float* out,
                          const float* g_out, float* g_vis, float* g_w, void* stream) {
  if (n == 0) return 0;
  spec_reduce_bwd_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(n, S, inv, rowB, vis, w, out, g_out, g_vis,
                                                                         g_w);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Shared sampling inputs of the two get_specular_visibility calls of render_with_all_sg (direct + indirect), doubled:
// ref / wl [2n][3], sharp [2n]; sg_range [1]; wlam [n] and argmin [1] are kept for the backward.
int robir_spec_prep_fwd(int n, const float* normal, const float* view, const float* rough, const unsigned char* valid,
                        float* ref, float* wl, float* sharp, float* sg_range, float* wlam, int* argmin, void* stream) {
  if (n == 0) return 0;
  spec_prep_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, normal, view, rough, valid, ref, wl, sharp, sg_range, wlam,
                                                             argmin);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_spec_prep_bwd(int n, const float* rough, const float* wlam, const int* argmin, const float* g_sharp,
                        const float* g_sg_range, float* g_rough, void* stream) {
  if (n == 0) return 0;
  spec_prep_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(n, rough, wlam, argmin, g_sharp, g_sg_range,
                                                                       g_rough);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
