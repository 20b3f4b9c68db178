// Tensor-core (tcgen05 / TMEM) engine for the fused visibility MLP, forward and input-gradient backward.
//
// One CTA per SM, persistent over 128-row tiles of the (point, direction) pair list.  Warp roles:
//   warp 0  : weight producer -- 1-D bulk async copies (UBLKCP) of pre-swizzled bf16 weight images from L2 into a
//             6-stage shared-memory ring (32 KB/stage = B_hi | B_lo for a 128(n) x 64(k) block, SWIZZLE_128B K-major);
//   warp 1  : MMA issuer -- one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16);
//   warps 2-5: epilogue -- one thread per row: tcgen05.ld the fp32 accumulators, bias + ReLU (fwd) / ReLU-mask (bwd),
//             split into bf16 hi/lo and tcgen05.st them back IN PLACE as the next layer's A operand.
// Activations never touch shared or global memory: the A operand of every layer lives in TMEM (two 256-column halves
// used ping-pong as "A of this layer" / "D of this layer"), so the whole 227 KB of shared memory feeds weights.
// fp32 parity: every product is evaluated as hi*hi + lo*hi + hi*lo in bf16 with fp32 accumulation (3 MMAs per
// logical one); measured error of the rendered colours vs. fp32 is ~1e-5 relative (hard part 5).
//
// TMEM layout of an A half (256 K-values of 128 rows in 256 columns): 32-column chunk c' holds K in [32c', 32c'+32):
//   columns [32c', 32c'+16) = packed bf16 hi (2 K-values per column), [32c'+16, 32c'+32) = packed bf16 lo,
// which is exactly where the epilogue finds the fp32 accumulators of the 32 output features it converts.
#include "common.cuh"
#include "tc_common.cuh"

namespace robir {
using namespace tc;

constexpr int kTcStages = 6;
constexpr int kTcStageBytes = 32768;
constexpr int kTcThreads = 192;
constexpr uint32_t kIdescN128 = idesc_bf16(128, 128);
constexpr uint32_t kIdescN64 = idesc_bf16(128, 64);

// ------------------------------------------------------------------------------------------------------------------
// weight image: for each (nh, kb): [hi: 128 rows x 64 k][lo: 128 rows x 64 k] bf16, SWIZZLE_128B (16-byte chunk c of
// row r stored at chunk c ^ (r & 7)).  B[n][k] = transpose ? W[k][n] : W[n][k];  rows >= N or k >= K are zero.
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_tc_image_kernel(const float* __restrict__ W, int ldw, int N, int K, int transpose,
                                     int n_halves, uint8_t* __restrict__ img) {
  // one thread per 16-byte chunk: 8 consecutive k of one row, hi and lo
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n_halves * 4 * 128 * 8;
  if (idx >= total) return;
  const int chunk = idx & 7, r = (idx >> 3) & 127, kb = (idx >> 10) & 3, nh = idx >> 12;
  const int n = nh * 128 + r;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = kb * 64 + chunk * 8 + 2 * j + e;
      float v = 0.f;
      if (n < N && k < K) v = transpose ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
      x[e] = v;
    }
    split_pack(x[0], x[1], hi[j], lo[j]);
  }
  uint8_t* stage = img + (size_t)(nh * 4 + kb) * kTcStageBytes;
  const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(stage + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(stage + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct TcParams {
  // pair list (128-row tiles)
  const int* rowA; const int* rowB; const int* n_tiles;
  // forward
  const float* tabA; const float* tabB;
  const uint8_t* img;      // fwd: [3 layers][2 nh][4 kb][32 KB]; bwd: [3 layers][2][4][32 KB] + [1][4][32 KB] (W0d, 64 rows)
  const float* bias;       // [3][256] (fwd)
  const float* wd;         // [256]
  const float* bd;         // [1]
  float* vis;              // [rows]
  uint32_t* mask;          // [tiles][4 layers][8 words][128 rows] (a warp stores one 128-byte line); bit i of word
                           // w of layer l, bit (31 - i) = (pre-activation of h_{l+1}[32 w + i] is not negative)
  // backward
  const float* g_vis; const float* dirs; float* g_dirs;
  // self-test: plain GEMM D = A . W^T through the same machinery
  const float* test_A; float* test_D; int test_mode;
};

template <int MODE>  // 0 = forward, 1 = backward, 2 = self-test (one 256x256 layer)
__global__ void __launch_bounds__(kTcThreads, 1) vis_tc_kernel(TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kTcStages], empty_bar[kTcStages], a_ready[2], d_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_a[128], s_b[128];
  __shared__ __align__(16) float s_bias[3 * 256];
  __shared__ __align__(16) float s_wd[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kLayers = (MODE == 0) ? 3 : (MODE == 1 ? 4 : 1);

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&a_ready[0], 128); mbar_init(&a_ready[1], 128);
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    fence_barrier_init();
  }
  if (MODE == 0) for (int i = tid; i < 3 * 256; i += kTcThreads) s_bias[i] = p.bias[i];
  if (MODE <= 1) for (int i = tid; i < 256; i += kTcThreads) s_wd[i] = p.wd[i];
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int ntiles = (MODE >= 2) ? 1 : *p.n_tiles;

  if (warp == 0) {
    // ===================================== weight producer =====================================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int layer = 0; layer < kLayers; ++layer) {
        const int halves = (MODE == 1 && layer == 3) ? 1 : 2;
        for (int nh = 0; nh < halves; ++nh)
          for (int kb = 0; kb < 4; ++kb, ++it) {
            const int st = it % kTcStages;
            mbar_wait(&empty_bar[st], ((it / kTcStages) & 1) ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full_bar[st], kTcStageBytes);
              bulk_g2s(ring + (size_t)st * kTcStageBytes, p.img + (size_t)((layer * 2 + nh) * 4 + kb) * kTcStageBytes,
                       kTcStageBytes, &full_bar[st]);
            }
            __syncwarp();
          }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // The whole warp runs the loop converged and one elected lane issues (elect.sync): tcgen05.mma / commit are
    // warp-uniform instructions, and inside a divergent `if (lane == 0)` ptxas wraps every one of them in an
    // ELECT / BRA.U.ANY retry loop (~75 issue cycles per MMA).
    uint32_t it = 0, a_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int layer = 0; layer < kLayers; ++layer, ++a_phase) {
        const uint32_t a_half = tmem_base + ((layer & 1) ? 256u : 0u);
        const uint32_t d_half = tmem_base + ((layer & 1) ? 0u : 256u);
        const bool last64 = (MODE == 1 && layer == 3);
        const int halves = last64 ? 1 : 2;
        for (int nh = 0; nh < halves; ++nh) {
          for (int kb = 0; kb < 4; ++kb, ++it) {
            if (nh == 0 && (kb == 0 || kb == 2)) mbar_wait(&a_ready[kb >> 1], a_phase & 1);   // A K-half available
            const int st = it % kTcStages;
            mbar_wait(&full_bar[st], (it / kTcStages) & 1);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint8_t* sb = ring + (size_t)st * kTcStageBytes;
              const uint64_t b_hi = smem_desc_sw128(sb), b_lo = smem_desc_sw128(sb + 16384);
              const uint32_t d_addr = d_half + 128u * nh;
              const uint32_t idesc = last64 ? kIdescN64 : kIdescN128;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int s = kb * 4 + j;                   // k16 step 0..15
                const uint32_t a_hi = a_half + 32u * (s >> 1) + 8u * (s & 1), a_lo = a_hi + 16u;
                if (kb == 0 && j == 0) umma_ts<0>(d_addr, a_hi, b_hi + 2u * j, idesc);
                else umma_ts<1>(d_addr, a_hi, b_hi + 2u * j, idesc);
                umma_ts<1>(d_addr, a_lo, b_hi + 2u * j, idesc);
                umma_ts<1>(d_addr, a_hi, b_lo + 2u * j, idesc);
              }
              umma_commit(&empty_bar[st]);
              if (kb == 3) umma_commit(&d_full[nh]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===================================== epilogue warps =====================================
    // Everything this role reads from global memory is requested one step before it is needed (next tile's row
    // indices and first two gather chunks during the last layer, gather chunk c + 2 while chunk c converts, the
    // backward's mask words one layer ahead): a single warp per scheduler cannot hide an L2 round trip otherwise.
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t d_phase[2] = {0, 0};
    // backward: the 8 ReLU-mask words the next consumer (stage 0 / layer epilogue) needs
    uint32_t mw[8];
    auto load_mask = [&](int t, int slot) {
#pragma unroll
      for (int w = 0; w < 8; ++w) mw[w] = __ldg(p.mask + ((size_t)t * 32 + slot * 8 + w) * 128 + row);
    };
    auto pick_mask = [&](int w) {      // register file has no dynamic indexing: 8 selects
      uint32_t v = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) v = (i == w) ? mw[i] : v;
      return v;
    };
    // per-tile inputs of this row
    int a_idx = 0, b_idx = -1;
    float g0 = 0.f;
    auto load_row = [&](int t, int& a, int& b, float& g) {
      a = p.rowA ? p.rowA[t * 128 + row] : 0;
      b = p.rowB[t * 128 + row];
      g = 0.f;
      if (MODE == 1 && b >= 0) {
        const float v = p.vis[t * 128 + row];
        g = p.g_vis[t * 128 + row] * v * (1.f - v);
      }
    };
    // forward: gather pipeline of layer-0 pre-activation chunks (32 features of tabA[a] and tabB[b] each)
    float4 ga[2][8], gb[2][8];
    auto gather = [&](int c, int slot, int a, int b) {
      const float4* pa = reinterpret_cast<const float4*>(p.tabA + (size_t)a * 256 + 32 * c);
      const float4* pb = reinterpret_cast<const float4*>(p.tabB + (size_t)(b >= 0 ? b : 0) * 256 + 32 * c);
#pragma unroll
      for (int i = 0; i < 8; ++i) { ga[slot][i] = __ldg(pa + i); gb[slot][i] = __ldg(pb + i); }
    };
    if (MODE <= 1 && blockIdx.x < ntiles) {
      load_row(blockIdx.x, a_idx, b_idx, g0);
      if (MODE == 1) load_mask(blockIdx.x, 3);
      if (MODE == 0) { gather(0, 0, a_idx, b_idx); gather(1, 1, a_idx, b_idx); }
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int q0 = tile * 128;
      const bool valid = b_idx >= 0;
      const bool has_next = tile + (int)gridDim.x < ntiles;
      uint32_t* mtile = p.mask ? p.mask + (size_t)tile * (32 * 128) + row : nullptr;
      // ---------------- stage 0: first A operand into TMEM half X (columns 0..255)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float x[32];
        if (MODE == 0) {
          uint32_t m = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 va = ga[c & 1][i], vb = gb[c & 1][i];
            const float v0 = va.x + vb.x, v1 = va.y + vb.y, v2 = va.z + vb.z, v3 = va.w + vb.w;
            m = __funnelshift_l(__float_as_uint(v0), m, 1);
            m = __funnelshift_l(__float_as_uint(v1), m, 1);
            m = __funnelshift_l(__float_as_uint(v2), m, 1);
            m = __funnelshift_l(__float_as_uint(v3), m, 1);
            x[4 * i] = fmaxf(v0, 0.f); x[4 * i + 1] = fmaxf(v1, 0.f);
            x[4 * i + 2] = fmaxf(v2, 0.f); x[4 * i + 3] = fmaxf(v3, 0.f);
          }
          if (c + 2 < 8) gather(c + 2, c & 1, a_idx, b_idx);
          if (mtile != nullptr && valid) mtile[(0 * 8 + c) * 128] = ~m;
        } else if (MODE == 1) {
          const uint32_t mword = valid ? mw[c] : 0u;
          const float4* wp = reinterpret_cast<const float4*>(s_wd + 32 * c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w = wp[i];
            x[4 * i] = ((int)(mword << (4 * i)) < 0) ? g0 * w.x : 0.f;
            x[4 * i + 1] = ((int)(mword << (4 * i + 1)) < 0) ? g0 * w.y : 0.f;
            x[4 * i + 2] = ((int)(mword << (4 * i + 2)) < 0) ? g0 * w.z : 0.f;
            x[4 * i + 3] = ((int)(mword << (4 * i + 3)) < 0) ? g0 * w.w : 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = p.test_A[(size_t)row * 256 + 32 * c + i];
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pack(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
        tmem_st16(tmem_base + lane_addr + 32u * c, hi);
        tmem_st16(tmem_base + lane_addr + 32u * c + 16u, lo);
        if (c == 3 || c == 7) {                       // K-half c / 4 of the first A operand is complete
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&a_ready[c >> 2]);
        }
      }
      if (MODE == 1) load_mask(tile, 2);
      int a_nxt = 0, b_nxt = -1;
      float g_nxt = 0.f;
      // ---------------- layers
      float logit = 0.f;
#pragma unroll 1
      for (int layer = 0; layer < kLayers; ++layer) {
        const uint32_t d_half = tmem_base + ((layer & 1) ? 0u : 256u);
        const bool last64 = (MODE == 1 && layer == 3);
        const int halves = last64 ? 1 : 2;
        if (MODE <= 1 && layer == kLayers - 1 && has_next) {       // next tile: rows, first gather chunks / mask words
          load_row(tile + gridDim.x, a_nxt, b_nxt, g_nxt);
          if (MODE == 1) load_mask(tile + gridDim.x, 3);
          if (MODE == 0) { gather(0, 0, a_nxt, b_nxt); gather(1, 1, a_nxt, b_nxt); }
        }
#pragma unroll 1
        for (int h = 0; h < halves; ++h) {
          mbar_wait(&d_full[h], d_phase[h] & 1);
          ++d_phase[h];
          tc_fence_after();
          if (last64) {
            // ---- backward tail: dPE[0..63] -> d dir via the PE jacobian
            uint32_t r0[32], r1[32];
            tmem_ld32(d_half + lane_addr, r0);
            tmem_ld32(d_half + lane_addr + 32u, r1);
            tmem_wait_ld();
            if (valid) {
              float dpe[64];
#pragma unroll
              for (int i = 0; i < 32; ++i) { dpe[i] = __uint_as_float(r0[i]); dpe[32 + i] = __uint_as_float(r1[i]); }
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float xin = __ldg(p.dirs + 3 * b_idx + i);
                float g = dpe[i], f = 1.f;
#pragma unroll
                for (int l = 0; l < 10; ++l) {
                  float sn, cs;
                  sincosf(xin * f, &sn, &cs);
                  g += f * (cs * dpe[3 + 6 * l + i] - sn * dpe[6 + 6 * l + i]);
                  f *= 2.f;
                }
                atomicAdd(p.g_dirs + 3 * b_idx + i, g);
              }
            }
          } else {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              const int n0 = 128 * h + 32 * c;                  // first output feature of this chunk
              const uint32_t taddr = d_half + lane_addr + (uint32_t)n0;
              uint32_t r[32];
              tmem_ld32(taddr, r);
              tmem_wait_ld();
              float x[32];
              if (MODE == 0) {
                const float4* bp = reinterpret_cast<const float4*>(s_bias + layer * 256 + n0);
                uint32_t m = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 bb = bp[i];
                  const float v0 = __uint_as_float(r[4 * i]) + bb.x, v1 = __uint_as_float(r[4 * i + 1]) + bb.y;
                  const float v2 = __uint_as_float(r[4 * i + 2]) + bb.z, v3 = __uint_as_float(r[4 * i + 3]) + bb.w;
                  m = __funnelshift_l(__float_as_uint(v0), m, 1);
                  m = __funnelshift_l(__float_as_uint(v1), m, 1);
                  m = __funnelshift_l(__float_as_uint(v2), m, 1);
                  m = __funnelshift_l(__float_as_uint(v3), m, 1);
                  x[4 * i] = fmaxf(v0, 0.f); x[4 * i + 1] = fmaxf(v1, 0.f);
                  x[4 * i + 2] = fmaxf(v2, 0.f); x[4 * i + 3] = fmaxf(v3, 0.f);
                }
                if (mtile != nullptr && valid) mtile[((layer + 1) * 8 + (n0 >> 5)) * 128] = ~m;
              } else if (MODE == 1) {
                const uint32_t mword = valid ? pick_mask(4 * h + c) : 0u;
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = ((int)(mword << i) < 0) ? __uint_as_float(r[i]) : 0.f;
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) p.test_D[(size_t)row * 256 + n0 + i] = __uint_as_float(r[i]);
              }
              if (MODE == 0 && layer == 2) {
                const float4* wp = reinterpret_cast<const float4*>(s_wd + n0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 w = wp[i];
                  logit = fmaf(w.x, x[4 * i], logit); logit = fmaf(w.y, x[4 * i + 1], logit);
                  logit = fmaf(w.z, x[4 * i + 2], logit); logit = fmaf(w.w, x[4 * i + 3], logit);
                }
              } else if (MODE <= 1) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_pack(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
                tmem_st16(taddr, hi);
                tmem_st16(taddr + 16u, lo);
              }
            }
            if (MODE <= 1 && !(MODE == 0 && layer == 2)) {
              tmem_wait_st();
              tc_fence_before();
              mbar_arrive(&a_ready[h]);
            }
          }
        }
        if (MODE == 1 && layer < 2) load_mask(tile, 1 - layer);     // words of the next layer's epilogue
      }
      if (MODE == 0) p.vis[q0 + row] = valid ? 1.f / (1.f + expf(-(logit + p.bd[0]))) : 0.f;
      a_idx = a_nxt; b_idx = b_nxt; g0 = g_nxt;
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace robir

using namespace robir;

static constexpr int kTcSmem = kTcStages * kTcStageBytes + 1024;

extern "C" {

// Packs one layer's weights into the tensor-core image.  transpose = 0: B[n][k] = W[n][k] (forward, W = torch weight
// [N][K]); transpose = 1: B[n][k] = W[k][n] (backward through the same layer).  n_halves = 2 for 256 rows, 1 for <= 128.
int robir_tc_pack_layer(const float* W, int ldw, int N, int K, int transpose, int n_halves, void* img, void* stream) {
  RB_REQUIRE(n_halves == 1 || n_halves == 2, "tc_pack_layer: n_halves must be 1 or 2");
  RB_REQUIRE(K <= 256 && N <= 128 * n_halves, "tc_pack_layer: shape exceeds the 256x256 image");
  const int total = n_halves * 4 * 128 * 8;
  pack_tc_image_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, transpose, n_halves,
                                                                               (uint8_t*)img);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_tc_image_bytes(int n_layers256, int n_layers64) { return (n_layers256 * 8 + n_layers64 * 8) * kTcStageBytes; }

// Forward over a 128-row-tile pair list.  img: 3 layers (robir_tc_pack_layer, transpose=0) back to back.
int robir_vis_tc_fwd(const float* tabA, const float* tabB, const int* rowA, const int* rowB, const int* n_tiles,
                     int max_tiles, const void* img, const float* bias3x256, const float* wd, const float* bd,
                     float* vis, uint32_t* mask, int sm_count, void* stream) {
  if (max_tiles == 0) return 0;
  TcParams p = {};
  p.rowA = rowA; p.rowB = rowB; p.n_tiles = n_tiles; p.tabA = tabA; p.tabB = tabB; p.img = (const uint8_t*)img;
  p.bias = bias3x256; p.wd = wd; p.bd = bd; p.vis = vis; p.mask = mask;
  RB_CHECK_CUDA(cudaFuncSetAttribute(vis_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  const int grid = max_tiles < sm_count ? max_tiles : sm_count;
  vis_tc_kernel<0><<<grid, kTcThreads, kTcSmem, (cudaStream_t)stream>>>(p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Backward (input gradient w.r.t. the direction).  img: W3^T, W2^T, W1^T (transpose=1, 2 halves each) then the
// 64-row W0d image (robir_tc_pack_layer(W0 + 63, ldw=126, N=63, K=256, transpose=1, n_halves=1)) in an 8-stage slot.
int robir_vis_tc_bwd(const int* rowB, const int* n_tiles, int max_tiles, const void* img, const float* wd,
                     const float* vis, const float* g_vis, const uint32_t* mask, const float* dirs, float* g_dirs,
                     int sm_count, void* stream) {
  if (max_tiles == 0) return 0;
  TcParams p = {};
  p.rowB = rowB; p.n_tiles = n_tiles; p.img = (const uint8_t*)img; p.wd = wd; p.vis = const_cast<float*>(vis);
  p.g_vis = g_vis; p.mask = const_cast<uint32_t*>(mask); p.dirs = dirs; p.g_dirs = g_dirs;
  RB_CHECK_CUDA(cudaFuncSetAttribute(vis_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  const int grid = max_tiles < sm_count ? max_tiles : sm_count;
  vis_tc_kernel<1><<<grid, kTcThreads, kTcSmem, (cudaStream_t)stream>>>(p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Self-test of the GEMM machinery: D[128][256] = A[128][256] . W[256][256]^T with the bf16x3 split (img packed with
// transpose=0, one layer).
int robir_tc_selftest(const float* A, const void* img, float* D, void* stream) {
  TcParams p = {};
  p.img = (const uint8_t*)img; p.test_A = A; p.test_D = D;
  RB_CHECK_CUDA(cudaFuncSetAttribute(vis_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  vis_tc_kernel<2><<<1, kTcThreads, kTcSmem, (cudaStream_t)stream>>>(p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
