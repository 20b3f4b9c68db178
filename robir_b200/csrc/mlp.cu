// Fused small-MLP kernels: the SparseAE encoders / decoders of the material network
// (model/sg_envmap_material.py:40-99) and the indirect-illumination lobe network
// (model/implicit_differentiable_renderer.py:170-222) run as ONE launch per chain instead of ~40 library kernels:
// input encoding (PE10 / PE10+hdr / IPE, optional additive noise in embedding space) -> up to 8 Linear(+ReLU|LeakyReLU)
// layers on 16-row tiles held in shared memory (fp32 FFMA, column-per-lane mapping, cp.async weight ring).  The backward kernel runs the
// input-gradient chain and emits the per-layer pre-activation gradients G_l; wgrad_kernel turns them into dW_l = G_l^T A_{l-1}
// and db_l (split over the rows, deterministic in-kernel reduction).
#include "mlp_engine.cuh"

namespace robir {

constexpr int kMlpMaxLayers = 8;
// rows per CTA tile: 16, or 8 when the batch is small (<= 2048 rows) so that a 1024-ray step still fills 128 SMs
constexpr int kMlpKMax = 512;

enum InMode { IN_RAW = 0, IN_PE10 = 1, IN_PE10_EXTRA = 2, IN_IPE10 = 3, IN_PE10X2 = 4 };

struct MlpLayer {
  const float* Wt;    // forward: [Kpad][Npad] (transposed, zero padded, Npad % 256 == 0, Kpad % 16 == 0)
  const float* Wb;    // backward: [Npad16][Kpad256] row-major copy of W (zero padded)
  const float* bias;  // [Npad]
  int K, N, Kpad, Npad;
  int act;            // Act of common.cuh applied to this layer's output
  float* save;        // forward: post-activation output [n][Npad] or null;  backward: same buffer (input)
  float* G;           // backward: pre-activation gradient of this layer [n][Npad] or null
};

struct MlpParams {
  int n, n_layers, in_mode, in_dim, in_pad;   // in_dim = K of layer 0; in_pad = Kpad of layer 0
  const float* x;       // IN_RAW: [n][in_dim]; else points [n][3]
  const float* extra;   // IN_PE10_EXTRA: [n] appended as column 63
  const float* noise;   // optional [n][in_dim], added as 0.02 * noise in embedding space (sg_envmap_material.py:83)
  float noise_scale;
  float* x0_save;       // embedded input [n][in_pad] (forward: written if non-null; backward: unused)
  MlpLayer L[kMlpMaxLayers];
  float* out;           // [n][ldo] first N_last columns written
  int ldo;
  const float* g_out;   // backward: [n][ldo]
  float* g_x;           // backward: gradient w.r.t. the embedded input [n][in_pad] or null
  const int* n_active;  // optional device scalar: only rows with (row % seg) < *n_active are evaluated (hit rays
                        // compacted to the front of a fixed-capacity batch); tiles without such a row write zeros
  int seg;              // rows per segment (the batch may be a concatenation of equally ordered copies); 0 = n
};

struct MlpActive { int n_act, seg; };
__device__ __forceinline__ MlpActive mlp_active(const MlpParams& p) {
  MlpActive a;
  a.seg = p.seg > 0 ? p.seg : (p.n > 0 ? p.n : 1);
  a.n_act = p.n_active ? min(__ldg(p.n_active), a.seg) : a.seg;
  return a;
}
__device__ __forceinline__ bool mlp_row_active(int row, const MlpActive& a) { return (row % a.seg) < a.n_act; }
// whole tile inside the inactive tail of one segment
__device__ __forceinline__ bool mlp_tile_inactive(int row0, int R, const MlpActive& a) {
  const int o = row0 % a.seg;
  return o >= a.n_act && o + R <= a.seg;
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_LEAKY02) return x > 0.f ? x : 0.2f * x;
  return x;
}
// derivative from the post-activation value (sign is preserved by ReLU / LeakyReLU)
__device__ __forceinline__ float act_bwd(float y, int act) {
  if (act == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == ACT_LEAKY02) return y > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// safe_trig_helper of model/neus_model.py:15-16: arguments with |a| >= 100 pi are wrapped with python-style %
__device__ __forceinline__ float safe_arg(float a) {
  const float T = 314.15927f;   // float32(100 * pi)
  if (fabsf(a) < T) return a;
  float m = fmodf(a, T);
  if (m != 0.f && m < 0.f) m += T;
  return m;
}

__device__ __forceinline__ void encode_row(float* Xs, int RP, int r, const MlpParams& p, int row, bool valid) {
  const int K = p.in_dim;
  if (p.in_mode == IN_RAW) {
    for (int k = 0; k < K; ++k) Xs[k * RP + r] = valid ? p.x[(size_t)row * K + k] : 0.f;
  } else if (p.in_mode == IN_PE10X2) {
    // VisNetwork input: [PE10(point) | PE10(direction)] from x = [n][6]  (implicit_differentiable_renderer.py:250-256)
    for (int h = 0; h < 2; ++h) {
      float v[3] = {0.f, 0.f, 0.f};
      if (valid) { v[0] = p.x[6 * row + 3 * h]; v[1] = p.x[6 * row + 3 * h + 1]; v[2] = p.x[6 * row + 3 * h + 2]; }
      const int o = 63 * h;
      for (int i = 0; i < 3; ++i) Xs[(o + i) * RP + r] = v[i];
      float f = 1.f;
      for (int l = 0; l < 10; ++l) {
        for (int i = 0; i < 3; ++i) {
          Xs[(o + 3 + 6 * l + i) * RP + r] = valid ? sinf(v[i] * f) : 0.f;
          Xs[(o + 6 + 6 * l + i) * RP + r] = valid ? cosf(v[i] * f) : 0.f;
        }
        f *= 2.f;
      }
    }
  } else {
    float v[3] = {0.f, 0.f, 0.f};
    if (valid) { v[0] = p.x[3 * row]; v[1] = p.x[3 * row + 1]; v[2] = p.x[3 * row + 2]; }
    if (p.in_mode == IN_IPE10) {
      // integrated PE, isotropic var 1e-5: [exp(-var 4^l / 2) sin(2^l x)] (30) then the same with +pi/2 (30)
      float f = 1.f;
      for (int l = 0; l < 10; ++l) {
        const float damp = expf(-0.5f * (1e-5f * f * f));
        for (int i = 0; i < 3; ++i) {
          const float y = v[i] * f;
          Xs[(3 * l + i) * RP + r] = valid ? damp * sinf(safe_arg(y)) : 0.f;
          Xs[(30 + 3 * l + i) * RP + r] = valid ? damp * sinf(safe_arg(y + 1.57079632679489661923f)) : 0.f;
        }
        f *= 2.f;
      }
    } else {
      for (int i = 0; i < 3; ++i) Xs[i * RP + r] = v[i];
      float f = 1.f;
      for (int l = 0; l < 10; ++l) {
        for (int i = 0; i < 3; ++i) {
          Xs[(3 + 6 * l + i) * RP + r] = valid ? sinf(v[i] * f) : 0.f;
          Xs[(6 + 6 * l + i) * RP + r] = valid ? cosf(v[i] * f) : 0.f;
        }
        f *= 2.f;
      }
      if (p.in_mode == IN_PE10_EXTRA) Xs[63 * RP + r] = valid ? p.extra[row] : 0.f;
    }
  }
  if (p.noise != nullptr && valid)
    for (int k = 0; k < K; ++k) Xs[k * RP + r] += p.noise_scale * p.noise[(size_t)row * K + k];
  for (int k = K; k < p.in_pad; ++k) Xs[k * RP + r] = 0.f;
  if (p.x0_save != nullptr && valid)
    for (int k = 0; k < p.in_pad; ++k) p.x0_save[(size_t)row * p.in_pad + k] = Xs[k * RP + r];
}

// Column-per-lane tile GEMM for narrow row tiles: warp w / lane l own output column 32 w + l of the current 256-column
// pass and all R rows of the tile (R accumulators in registers); per k one conflict-free LDS.32 of the weight row and
// R/4 broadcast LDS.128 of the activations -> FFMA-issue bound even at R = 16, which lets 1024 rows spread over 64 CTAs.
template <int R>
__device__ __forceinline__ void col_gemm_pass(const float* __restrict__ Xs, int K, const float* __restrict__ Wt,
                                              int ldw, int col0, float* __restrict__ Wbuf, float (&acc)[R]) {
  constexpr int RP = R + 4;
  const int tid = threadIdx.x;
  auto load_chunk = [&](int buf, int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 6, c4 = idx & 63;
      cp_async16(Wbuf + buf * (kChunkK * kPassCols) + row * kPassCols + c4 * 4,
                 Wt + (size_t)(k0 + row) * ldw + col0 + c4 * 4);
    }
    cp_async_commit();
  };
  const int nchunk = K / kChunkK;
  load_chunk(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) {
      load_chunk((c + 1) & 1, (c + 1) * kChunkK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = Wbuf + (c & 1) * (kChunkK * kPassCols) + tid;
    const float* xb = Xs + (size_t)(c * kChunkK) * RP;
#pragma unroll
    for (int kk = 0; kk < kChunkK; ++kk) {
      const float w = wb[kk * kPassCols];
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 x = *reinterpret_cast<const float4*>(xb + kk * RP + r);
        acc[r] = fmaf(x.x, w, acc[r]);
        acc[r + 1] = fmaf(x.y, w, acc[r + 1]);
        acc[r + 2] = fmaf(x.z, w, acc[r + 2]);
        acc[r + 3] = fmaf(x.w, w, acc[r + 3]);
      }
    }
    __syncthreads();
  }
}

template <int R>
__device__ __forceinline__ void col_store(float* __restrict__ Xs, int col, const float (&v)[R]) {
  float* dst = Xs + (size_t)col * (R + 4);
#pragma unroll
  for (int r = 0; r < R; r += 4) *reinterpret_cast<float4*>(dst + r) = make_float4(v[r], v[r + 1], v[r + 2], v[r + 3]);
}

template <int R, int NPASS>
__device__ __forceinline__ void mlp_fwd_layer(const MlpParams& p, int l, float* Xs, float* Wbuf, int row0) {
  const MlpLayer& L = p.L[l];
  const bool last = l == p.n_layers - 1;
  float acc[NPASS][R];
#pragma unroll
  for (int ps = 0; ps < NPASS; ++ps) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[ps][r] = 0.f;
    col_gemm_pass<R>(Xs, L.Kpad, L.Wt, L.Npad, ps * kPassCols, Wbuf, acc[ps]);
  }
#pragma unroll
  for (int ps = 0; ps < NPASS; ++ps) {
    const int col = ps * kPassCols + threadIdx.x;
    const bool live = col < L.N;
    const float b = live ? __ldg(L.bias + col) : 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float y = live ? act_fwd(acc[ps][r] + b, L.act) : 0.f;
      acc[ps][r] = y;
      const int row = row0 + r;
      if (row < p.n) {
        if (L.save != nullptr) L.save[(size_t)row * L.Npad + col] = y;
        if (last && live) p.out[(size_t)row * p.ldo + col] = y;
      }
    }
    if (!last) col_store<R>(Xs, col, acc[ps]);
  }
  __syncthreads();
}

template <int R>
__global__ void __launch_bounds__(256, 2) mlp_fwd_kernel(MlpParams p) {
  constexpr int RP = R + 4;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                        // [kMlpKMax][RP]
  float* Wbuf = Xs + kMlpKMax * RP;
  const int tid = threadIdx.x;
  const int ntile = (p.n + R - 1) / R;
  const MlpActive act = mlp_active(p);
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int row0 = tile * R;
    if (mlp_tile_inactive(row0, R, act)) {  // zero every output row, no arithmetic
      const int rows = min(R, p.n - row0);
      for (int l = 0; l < p.n_layers; ++l)
        if (p.L[l].save != nullptr)
          for (int i = tid; i < rows * p.L[l].Npad; i += 256) p.L[l].save[(size_t)row0 * p.L[l].Npad + i] = 0.f;
      if (p.x0_save != nullptr)
        for (int i = tid; i < rows * p.in_pad; i += 256) p.x0_save[(size_t)row0 * p.in_pad + i] = 0.f;
      for (int i = tid; i < rows * p.ldo; i += 256) p.out[(size_t)row0 * p.ldo + i] = 0.f;
      continue;
    }
    __syncthreads();
    if (tid < R) encode_row(Xs, RP, tid, p, row0 + tid, row0 + tid < p.n && mlp_row_active(row0 + tid, act));
    __syncthreads();
    for (int l = 0; l < p.n_layers; ++l) {
      if (p.L[l].Npad == kPassCols) mlp_fwd_layer<R, 1>(p, l, Xs, Wbuf, row0);
      else mlp_fwd_layer<R, 2>(p, l, Xs, Wbuf, row0);
    }
  }
}

template <int R, int NPASS>
__device__ __forceinline__ void mlp_bwd_layer(const MlpParams& p, int l, float* Xs, float* Wbuf, int row0) {
  const MlpLayer& L = p.L[l];
  const int kin = (L.N + 15) & ~15;                 // contraction length (rows of Wb)
  const int outp = NPASS * kPassCols;
  float acc[NPASS][R];
#pragma unroll
  for (int ps = 0; ps < NPASS; ++ps) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[ps][r] = 0.f;
    col_gemm_pass<R>(Xs, kin, L.Wb, outp, ps * kPassCols, Wbuf, acc[ps]);
  }
#pragma unroll
  for (int ps = 0; ps < NPASS; ++ps) {
    const int col = ps * kPassCols + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r;
      float g = (col < L.K) ? acc[ps][r] : 0.f;
      if (l > 0) {
        const MlpLayer& P = p.L[l - 1];
        if (row < p.n && col < P.N) {
          g *= act_bwd(P.save[(size_t)row * P.Npad + col], P.act);
          if (P.G != nullptr) P.G[(size_t)row * P.Npad + col] = g;
        } else {
          g = 0.f;
        }
      } else if (p.g_x != nullptr && row < p.n && col < p.in_pad) {
        p.g_x[(size_t)row * p.in_pad + col] = g;
      }
      acc[ps][r] = g;
    }
    if (l > 0) col_store<R>(Xs, col, acc[ps]);
  }
  __syncthreads();
}

// backward: G_L = g_out; for l = L-1..0: emit G_l, dA = G_l . W_l, G_{l-1} = dA * act'(A_{l-1})
template <int R>
__global__ void __launch_bounds__(256, 2) mlp_bwd_kernel(MlpParams p) {
  constexpr int RP = R + 4;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Wbuf = Xs + kMlpKMax * RP;
  const int tid = threadIdx.x;
  const int ntile = (p.n + R - 1) / R;
  const MlpActive act = mlp_active(p);
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int row0 = tile * R;
    if (mlp_tile_inactive(row0, R, act)) {  // zero gradients
      const int rows = min(R, p.n - row0);
      for (int l = 0; l < p.n_layers; ++l)
        if (p.L[l].G != nullptr)
          for (int i = tid; i < rows * p.L[l].Npad; i += 256) p.L[l].G[(size_t)row0 * p.L[l].Npad + i] = 0.f;
      if (p.g_x != nullptr)
        for (int i = tid; i < rows * p.in_pad; i += 256) p.g_x[(size_t)row0 * p.in_pad + i] = 0.f;
      continue;
    }
    __syncthreads();
    {
      // G of the last layer (its act is applied by the caller or is NONE): tile rows <- g_out, zero padded to Npad16
      const MlpLayer& L = p.L[p.n_layers - 1];
      const int kp = (L.N + 15) & ~15;
      for (int idx = tid; idx < kp * R; idx += 256) {
        const int k = idx / R, r = idx % R, row = row0 + r;
        float g = 0.f;
        if (row < p.n && k < L.N) {
          g = p.g_out[(size_t)row * p.ldo + k];
          if (L.act != ACT_NONE) g *= act_bwd(L.save[(size_t)row * L.Npad + k], L.act);
        }
        Xs[k * RP + r] = g;
        if (L.G != nullptr && row < p.n && k < L.N) L.G[(size_t)row * L.Npad + k] = g;
      }
    }
    __syncthreads();
    for (int l = p.n_layers - 1; l >= 0; --l) {
      if (p.L[l].K <= kPassCols) mlp_bwd_layer<R, 1>(p, l, Xs, Wbuf, row0);
      else mlp_bwd_layer<R, 2>(p, l, Xs, Wbuf, row0);
    }
  }
}

// Weight / bias gradient of one Linear layer from the pre-activation gradients G [n][ldg] and the layer input
// A [n][lda]:  dW[i][j] = sum_r G[r][i] A[r][j],  db[i] = sum_r G[r][i]   (i < N, j < K), fp32, fixed summation order.
// Rows are visited in chunks of 32; chunks that lie entirely in the inactive tail of a segment (n_active, see
// MlpParams) are skipped.  One CTA per (64 x 64 tile of dW, row split z of S): chunk c belongs to split c % S; the next
// chunk is prefetched into registers while the current one is multiplied.  With S > 1 every CTA writes its partial tile
// to a workspace and the last one to arrive (ticket counter, reset for the next launch) adds the S partials in split
// order -- a deterministic split-K without a second launch.
constexpr int kWgRows = 32;
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ A,
                                                    int lda, int n, int N, int K, const int* __restrict__ n_active,
                                                    int seg, float* __restrict__ partial, int* __restrict__ tickets,
                                                    float* __restrict__ dW, float* __restrict__ db) {
  __shared__ __align__(16) float Gs[kWgRows][64];
  __shared__ __align__(16) float As[kWgRows][64];
  __shared__ int s_last;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64, z = blockIdx.z, S = gridDim.z;
  const int sg = seg > 0 ? seg : (n > 0 ? n : 1);
  const int n_act = n_active ? min(__ldg(n_active), sg) : sg;
  const int nchunk = (n + kWgRows - 1) / kWgRows;
  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  auto active = [&](int c) {
    const int o = (c * kWgRows) % sg;
    return !(o >= n_act && o + kWgRows <= sg);
  };
  auto next_chunk = [&](int c) {          // next active chunk of this split at or after c (c % S == z)
    while (c < nchunk && !active(c)) c += S;
    return c;
  };
  // each thread stages 8 G and 8 A values per chunk: rows lr, lr + 4, ...  (column = tid & 63)
  const int lc = tid & 63, lr = tid >> 6;
  float rg[8], ra[8];
  auto fetch = [&](int c) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int row = c * kWgRows + lr + 4 * q;
      rg[q] = (row < n && i0 + lc < N) ? __ldg(G + (size_t)row * ldg + i0 + lc) : 0.f;
      ra[q] = (row < n && j0 + lc < K) ? __ldg(A + (size_t)row * lda + j0 + lc) : 0.f;
    }
  };
  int c = next_chunk(z);
  if (c < nchunk) fetch(c);
  while (c < nchunk) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) { Gs[lr + 4 * q][lc] = rg[q]; As[lr + 4 * q][lc] = ra[q]; }
    __syncthreads();
    c = next_chunk(c + S);
    if (c < nchunk) fetch(c);              // in flight during the multiply below
#pragma unroll 8
    for (int r = 0; r < kWgRows; ++r) {
      const float4 g = *reinterpret_cast<const float4*>(&Gs[r][ty * 4]);
      const float4 a = *reinterpret_cast<const float4*>(&As[r][tx * 4]);
      const float gv[4] = {g.x, g.y, g.z, g.w}, av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        bsum[x] += gv[x];
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(gv[x], av[y], acc[x][y]);
      }
    }
  }
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (S > 1) {
    // partial layout: [z][tile][64 x 64 values | 64 bias sums]
    float* mine = partial + ((size_t)z * gridDim.x * gridDim.y + tile) * (64 * 64 + 64);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
#pragma unroll
      for (int y = 0; y < 4; ++y) mine[(ty * 4 + x) * 64 + tx * 4 + y] = acc[x][y];
      if (tx == 0) mine[64 * 64 + ty * 4 + x] = bsum[x];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int t = atomicAdd(&tickets[tile], 1);
      s_last = (t == S - 1);
      if (s_last) tickets[tile] = 0;       // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      bsum[x] = 0.f;
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    }
    for (int s = 0; s < S; ++s) {
      const float* src = partial + ((size_t)s * gridDim.x * gridDim.y + tile) * (64 * 64 + 64);
#pragma unroll
      for (int x = 0; x < 4; ++x) {
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] += __ldcg(src + (ty * 4 + x) * 64 + tx * 4 + y);
        if (tx == 0) bsum[x] += __ldcg(src + 64 * 64 + ty * 4 + x);
      }
    }
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int i = i0 + ty * 4 + x;
    if (i >= N) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int j = j0 + tx * 4 + y;
      if (j < K) dW[(size_t)i * K + j] = acc[x][y];
    }
    if (db != nullptr && blockIdx.y == 0 && tx == 0) db[i] = bsum[x];
  }
}

// Embedded network input alone (PE / IPE / raw + noise + extra column) -> x0_save [n][in_pad]; rows outside the active
// head are written as zeros.  Front end of the tensor-core layer engine (tc_mlp.cu).
__global__ void __launch_bounds__(256) mlp_encode_kernel(MlpParams p) {
  constexpr int R = 8, RP = R + 4;
  __shared__ float Xs[kMlpKMax / 4 * RP];        // in_pad <= 128 rows of the k-major tile
  const int tid = threadIdx.x;
  const MlpActive act = mlp_active(p);
  const int ntile = (p.n + R - 1) / R;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int row0 = tile * R;
    __syncthreads();
    if (tid < R) {
      MlpParams q = p;
      q.x0_save = nullptr;
      encode_row(Xs, RP, tid, q, row0 + tid, row0 + tid < p.n && mlp_row_active(row0 + tid, act));
    }
    __syncthreads();
    for (int i = tid; i < R * p.in_pad; i += 256) {
      const int r = i / p.in_pad, k = i % p.in_pad;
      if (row0 + r < p.n) p.x0_save[(size_t)(row0 + r) * p.in_pad + k] = Xs[k * RP + r];
    }
  }
}

// W [N][K] -> Wb [Npad16][Kpad256] (zero padded row-major copy for the backward chain)
__global__ void pack_pad_kernel(const float* __restrict__ W, int N, int K, float* __restrict__ out, int Np, int Kp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Np * Kp) return;
  const int n = idx / Kp, k = idx % Kp;
  out[idx] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
}

// every packed copy of a chain of plain Linear layers in ONE launch (the re-pack of the trained chains sits at the head of
// the step once the octree walk is pipelined away): per layer Wt [Kp][Np] = W^T zero padded, Wb [Nb][Kb] = W zero padded,
// bias [Np] = b zero padded; blockIdx.y = layer
struct PackChainLayer {
  const float* W; const float* b;
  float* Wt; float* Wb; float* bias;
  int N, K, Kp, Np, Nb, Kb;
};
struct PackChainParams {
  int n_layers;
  PackChainLayer L[8];
};
__global__ void pack_chain_kernel(PackChainParams p) {
  const PackChainLayer& l = p.L[blockIdx.y];
  const int n_wt = l.Kp * l.Np, n_wb = l.Nb * l.Kb;
  const int total = n_wt + n_wb + l.Np;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    if (idx < n_wt) {
      const int k = idx / l.Np, n = idx % l.Np;
      l.Wt[idx] = (k < l.K && n < l.N) ? l.W[(size_t)n * l.K + k] : 0.f;
    } else if (idx < n_wt + n_wb) {
      const int j = idx - n_wt, n = j / l.Kb, k = j % l.Kb;
      l.Wb[j] = (n < l.N && k < l.K) ? l.W[(size_t)n * l.K + k] : 0.f;
    } else {
      const int n = idx - n_wt - n_wb;
      l.bias[n] = n < l.N ? l.b[n] : 0.f;
    }
  }
}

}  // namespace robir

using namespace robir;

static int mlp_check(const MlpParams* p) {
  RB_REQUIRE(p->n_layers >= 1 && p->n_layers <= kMlpMaxLayers, "mlp: 1..8 layers");
  for (int l = 0; l < p->n_layers; ++l) {
    const MlpLayer& L = p->L[l];
    RB_REQUIRE(L.Kpad % 16 == 0 && L.Kpad <= kMlpKMax && L.Npad % 256 == 0 && L.Npad <= 512 && L.K <= L.Kpad &&
                   L.N <= L.Npad,
               "mlp: layer shape outside the supported envelope (K <= 512, N <= 512)");
    if (l > 0) RB_REQUIRE(L.Kpad == p->L[l - 1].Npad || L.K == p->L[l - 1].N, "mlp: layer chain mismatch");
  }
  RB_REQUIRE(p->in_pad == p->L[0].Kpad && p->in_dim == p->L[0].K, "mlp: input width mismatch");
  return 0;
}

template <bool FWD, int R>
static int mlp_launch_r(const MlpParams* p, int sm_count, void* stream) {
  const int smem = (kMlpKMax * (R + 4) + kWbufFloats) * 4;
  auto kern = FWD ? mlp_fwd_kernel<R> : mlp_bwd_kernel<R>;
  RB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int tiles = (p->n + R - 1) / R;
  kern<<<tiles < 2 * sm_count ? tiles : 2 * sm_count, 256, smem, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
template <bool FWD>
static int mlp_launch(const MlpParams* p, int sm_count, void* stream) {
  return p->n <= 2048 ? mlp_launch_r<FWD, 8>(p, sm_count, stream) : mlp_launch_r<FWD, 16>(p, sm_count, stream);
}

extern "C" {

int robir_pack_chain(const PackChainParams* p, void* stream) {
  RB_REQUIRE(p->n_layers >= 1 && p->n_layers <= 8, "pack_chain: 1..8 layers");
  for (int l = 0; l < p->n_layers; ++l)
    RB_REQUIRE(p->L[l].Kp >= p->L[l].K && p->L[l].Np >= p->L[l].N && p->L[l].Nb >= p->L[l].N && p->L[l].Kb >= p->L[l].K,
               "pack_chain: padded shape too small");
  pack_chain_kernel<<<dim3(64, p->n_layers), 256, 0, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_pack_pad(const float* W, int N, int K, float* out, int Np, int Kp, void* stream) {
  RB_REQUIRE(Np >= N && Kp >= K, "pack_pad: padded shape too small");
  pack_pad_kernel<<<(Np * Kp + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, N, K, out, Np, Kp);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_mlp_fwd(const MlpParams* p, int sm_count, void* stream) {
  if (p->n == 0) return 0;
  if (int e = mlp_check(p)) return e;
  return mlp_launch<true>(p, sm_count, stream);
}

int robir_mlp_bwd(const MlpParams* p, int sm_count, void* stream) {
  if (p->n == 0) return 0;
  if (int e = mlp_check(p)) return e;
  for (int l = 0; l < p->n_layers - 1; ++l)
    RB_REQUIRE(p->L[l].save != nullptr, "mlp_bwd: hidden activations must have been saved by the forward");
  return mlp_launch<false>(p, sm_count, stream);
}

// dW [N][K] = G[:, :N]^T A[:, :K], db [N] = column sums of G (db may be null); G / A row strides ldg / lda (floats).
// splits > 1: rows are divided over that many CTAs per tile; workspace (caller-owned): partial
// [splits * tiles * 4160] floats and tickets [tiles] int32 (zero before the first call; the kernel re-zeroes them),
// tiles = ceil(N / 64) * ceil(K / 64).
int robir_mlp_wgrad(const float* G, int ldg, const float* A, int lda, int n, int N, int K, const int* n_active, int seg,
                    int splits, float* partial, int* tickets, float* dW, float* db, void* stream) {
  if (N == 0 || K == 0) return 0;
  RB_REQUIRE(seg == 0 || seg % kWgRows == 0 || n_active == nullptr, "mlp_wgrad: segment length must be a multiple of 32");
  RB_REQUIRE(splits >= 1 && splits <= 64 && (splits == 1 || (partial != nullptr && tickets != nullptr)),
             "mlp_wgrad: 1..64 splits, workspace required for splits > 1");
  dim3 grid((N + 63) / 64, (K + 63) / 64, splits);
  wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(G, ldg, A, lda, n, N, K, n_active, seg, partial, tickets, dW, db);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// x0_save [n][in_pad] = embedded input of the chain (only n, in_mode, in_dim, in_pad, x, extra, noise, noise_scale,
// n_active, seg and x0_save of the parameter block are read)
int robir_mlp_encode(const MlpParams* p, int sm_count, void* stream) {
  if (p->n == 0) return 0;
  RB_REQUIRE(p->x0_save != nullptr && p->in_pad <= kMlpKMax / 4 && p->in_dim <= p->in_pad, "mlp_encode: bad input block");
  const int tiles = (p->n + 7) / 8;
  mlp_encode_kernel<<<tiles < 4 * sm_count ? tiles : 4 * sm_count, 256, 0, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
