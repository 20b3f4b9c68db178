// NeuS SDF network evaluation: value, input gradient (normal) and 256-d feature in one fused
// kernel.  Reference: model/neus_model.py:312-438 (SDFNetwork: PE(10) -> 9 weight-normed linears, skip at layer 4,
// Softplus(beta=100)), :785-818 (ImplicitNetworkMy.forward/gradient: f(p) = net(2p)/2).
// The gradient is propagated in forward mode (three tangent rows ride along each value row through the same tile
// GEMMs), so nothing is stored and no second pass is needed; the reference's autograd double-backward graph is not
// reproduced (it is never consumed: SURVEY.md section 7, hard part 6).
#include "sdf_tile.cuh"

namespace robir {

// folded weight-norm + transpose:  Wt[k][n] = g[n] * v[n][k] / ||v[n]||   for n in [n_begin, n_begin + n_count)
__global__ void pack_wn_transpose_kernel(const float* __restrict__ v, const float* __restrict__ g, int N, int K,
                                         int n_begin, int n_count, float* __restrict__ Wt, int Kpad, int Npad) {
  // one warp per destination column n
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= Npad) return;
  const bool ok = n < n_count;
  const float* row = v + (size_t)(n_begin + n) * K;
  float ss = 0.f;
  if (ok)
    for (int k = lane; k < K; k += 32) ss += row[k] * row[k];
  ss = warp_sum(ss);
  const float s = ok ? g[n_begin + n] / sqrtf(ss) : 0.f;
  for (int k = lane; k < Kpad; k += 32) Wt[(size_t)k * Npad + n] = (ok && k < K) ? row[k] * s : 0.f;
}
// folded weight-norm, single row kept row-major: out[k] = g[n] v[n][k] / ||v[n]||
__global__ void pack_wn_row_kernel(const float* __restrict__ v, const float* __restrict__ g, int K, int n,
                                   float* __restrict__ out) {
  const int lane = threadIdx.x;
  const float* row = v + (size_t)n * K;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) ss += row[k] * row[k];
  ss = warp_sum(ss);
  const float s = g[n] / sqrtf(ss);
  for (int k = lane; k < K; k += 32) out[k] = row[k] * s;
}

struct SdfParams {
  const float* pts;     // [n][3]
  int n;
  float in_scale;       // 2.0 for stage-2 points (ImplicitNetworkMy.normalize), 1.0 for NeuS coordinates
  float sdf_scale;      // 0.5 / 1.0
  float feat_scale;
  const float* Wt[8];   // layers 0..7 packed [Kpad][256] (layer 0: K=64; layer 3: 193 valid columns)
  const float* bias[8];
  const float* w8_sdf;  // [256] folded row 0 of layer 8
  const float* b8;      // [257]
  const float* Wt8_feat;  // [256][256] folded rows 1..256 of layer 8, transposed (may be null)
  float* sdf;           // [n]
  float* grad;          // [n][3] or null
  float* feat;          // [n][256] or null
  const int* n_active;  // optional device scalar: points at or beyond min(*n_active, n) are not evaluated (outputs = 0)
};

// R = tile rows (64, or 32 for small batches so that a 1024-point call still spreads over ~128 CTAs)
template <bool JET, int R>
__global__ void __launch_bounds__(256, 2) sdf_eval_kernel(SdfParams p) {
  constexpr int RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  constexpr int PTS = JET ? R / 4 : R;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Wbuf = Xs + 256 * RP;
  float* red = Wbuf + kWbufFloats;  // [256 / R][R]
  __shared__ float s_x[PTS][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntile = (p.n + PTS - 1) / PTS;
  const float kInvSqrt2 = 0.70710678118654752440f;
  const int n_act = p.n_active ? min(__ldg(p.n_active), p.n) : p.n;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int p0 = tile * PTS;
    if (p0 >= n_act) {                     // inactive tile (hit rays are compacted to the front): zero outputs
      const int pts = min(PTS, p.n - p0);
      for (int i = tid; i < pts; i += 256) p.sdf[p0 + i] = 0.f;
      if (p.grad) for (int i = tid; i < pts * 3; i += 256) p.grad[(size_t)p0 * 3 + i] = 0.f;
      if (p.feat) for (int i = tid; i < pts * 256; i += 256) p.feat[(size_t)p0 * 256 + i] = 0.f;
      continue;
    }
    __syncthreads();
    if (tid < PTS) {
      const int i = p0 + tid;
      for (int c = 0; c < 3; ++c) s_x[tid][c] = i < p.n ? p.pts[3 * i + c] * p.in_scale : 0.f;
    }
    __syncthreads();
    if (tid < PTS) {
      sdf_pe_rows<JET>(Xs, RP, 0, tid, s_x[tid], p0 + tid < p.n, 1.f);
      for (int r = 0; r < (JET ? 4 : 1); ++r) Xs[63 * RP + (JET ? tid * 4 + r : tid)] = 0.f;
    }
    __syncthreads();
    float acc[TR][8];
    for (int layer = 0; layer < 8; ++layer) {
      zero_acc<R>(acc);
      tile_gemm_pass<R>(Xs, layer == 0 ? 64 : 256, p.Wt[layer], 256, 0, Wbuf, acc);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias[layer] + lane * 4));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias[layer] + 128 + lane * 4));
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float post = layer == 3 ? kInvSqrt2 : 1.f;   // x = cat([x, pe]) / sqrt(2) before layer 4
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (JET) {
#pragma unroll
          for (int h = 0; h < TR; h += 4) {
            const float z = acc[h][c] + bb[c];
            const float d = softplus100_grad(z);
            acc[h][c] = softplus100(z) * post;
            acc[h + 1][c] *= d * post;
            acc[h + 2][c] *= d * post;
            acc[h + 3][c] *= d * post;
          }
        } else {
#pragma unroll
          for (int r = 0; r < TR; ++r) acc[r][c] = softplus100(acc[r][c] + bb[c]) * post;
        }
      }
      store_acc<R>(Xs, 0, layer == 3 ? 193 : 256, acc);
      if (layer == 3) {
        __syncthreads();
        if (tid < PTS) sdf_pe_rows<JET>(Xs, RP, 193, tid, s_x[tid], p0 + tid < p.n, kInvSqrt2);
      }
      __syncthreads();
    }
    // ---- layer 8, column 0 (sdf / gradient components): dot over the 256 activations of every row
    {
      constexpr int PARTS = 256 / R, KLEN = 256 / PARTS;
      const int row = tid % R, part = tid / R;
      float s = 0.f;
      for (int k = part * KLEN; k < part * KLEN + KLEN; ++k) s = fmaf(__ldg(p.w8_sdf + k), Xs[k * RP + row], s);
      red[part * R + row] = s;
    }
    __syncthreads();
    if (tid < R) {
      float d;
      if (R == 64) {
        d = (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
      } else {
        d = 0.f;
#pragma unroll
        for (int q = 0; q < 256 / R; ++q) d += red[q * R + tid];
      }
      if (JET) {
        const int i = p0 + (tid >> 2), j = tid & 3;
        if (i < p.n) {
          if (j == 0) p.sdf[i] = (d + __ldg(p.b8)) * p.sdf_scale;
          else if (p.grad) p.grad[3 * i + j - 1] = d;   // d f(p) / d p = d net / d x  (f = net(2p)/2)
        }
      } else if (p0 + tid < p.n) {
        p.sdf[p0 + tid] = (d + __ldg(p.b8)) * p.sdf_scale;
      }
    }
    // ---- layer 8, feature columns
    if (p.feat != nullptr) {
      zero_acc<R>(acc);
      tile_gemm_pass<R>(Xs, 256, p.Wt8_feat, 256, 0, Wbuf, acc);
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const int row = warp * TR + r;
        int i;
        if (JET) {
          if (row & 3) continue;
          i = p0 + (row >> 2);
        } else {
          i = p0 + row;
        }
        if (i >= p.n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = tile_col(lane, c);
          p.feat[(size_t)i * 256 + col] = (acc[r][c] + __ldg(p.b8 + 1 + col)) * p.feat_scale;
        }
      }
    }
  }
}

}  // namespace robir

using namespace robir;
static inline int cdiv_i(long long a, long long b) { return (int)((a + b - 1) / b); }

extern "C" {

int robir_pack_wn_transpose(const float* v, const float* g, int N, int K, int n_begin, int n_count, float* Wt,
                            int Kpad, int Npad, void* stream) {
  RB_REQUIRE(n_begin >= 0 && n_begin + n_count <= N && n_count <= Npad && K <= Kpad, "pack_wn_transpose: bad window");
  pack_wn_transpose_kernel<<<cdiv_i(Npad, 4), 128, 0, (cudaStream_t)stream>>>(v, g, N, K, n_begin, n_count, Wt, Kpad,
                                                                              Npad);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_pack_wn_row(const float* v, const float* g, int K, int n, float* out, void* stream) {
  pack_wn_row_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(v, g, K, n, out);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// SDF network: sdf [n] (always), grad [n][3] (optional -> jet mode), feat [n][256] (optional)
int robir_sdf_eval(const SdfParams* p, int sm_count, void* stream) {
  if (p->n == 0) return 0;
  const int smem = (256 * 68 + kWbufFloats + 256) * 4;
  const bool jet = p->grad != nullptr;
  RB_REQUIRE(p->feat == nullptr || p->Wt8_feat != nullptr, "sdf_eval: feature output needs Wt8_feat");
  const bool small = (long long)p->n * (jet ? 4 : 1) <= 4096;   // 32-row tiles: 1024 points -> 128 (jet) / 32 CTAs
  const int pts = (small ? 32 : 64) / (jet ? 4 : 1);
  const int tiles = cdiv_i(p->n, pts);
  const int grid = tiles < 2 * sm_count ? tiles : 2 * sm_count;
  void (*kern)(SdfParams) = jet ? (small ? sdf_eval_kernel<true, 32> : sdf_eval_kernel<true, 64>)
                                : (small ? sdf_eval_kernel<false, 32> : sdf_eval_kernel<false, 64>);
  RB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, 256, smem, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
