// NeuS SDF network on the tensor cores (tcgen05 / TMEM): value, forward-mode normal (three tangent rows ride along each
// value row) and optional 256-d feature of model/neus_model.py:312-438, 785-818 in one persistent kernel -- the same
// machine as the fused visibility MLP (csrc/vis_tc.cu): warp 0 streams pre-swizzled scaled-fp16 hi/lo weight images
// through a shared-memory ring, warp 1 issues M128.N128.K16 MMAs with the A operand in TMEM, eight epilogue warps
// convert the fp32 accumulators in place into the next layer's A operand.  fp32 parity: three MMAs per logical product
// (hi*hi + lo*hi + hi*lo), fp32 accumulation.
//
// Tile = 128 rows = 32 points x (value row, d/dx, d/dy, d/dz rows) in jet mode, 128 points otherwise.
// Layers: PE(10) 63 -> 256 (K = 64: one K block), 256 -> 256 x2, 256 -> 193, cat([h, PE]) / sqrt(2) (the 1/sqrt(2)
// is folded into the next layer's image) -> 256 x4, Softplus(beta = 100) after each; the epilogue applies the
// activation to the value rows and its derivative sigmoid(100 z) of the VALUE row (warp shuffle) to the tangent rows.
// Layer 8 (256 -> 257): column 0 (sdf / normal components) is an fp32 dot product in the last epilogue, the 256 feature
// columns are one more MMA layer when requested.
#include "common.cuh"
#include "tc_common.cuh"

namespace robir {
using namespace tc;

constexpr int kSdfRingBytes = 196608;
constexpr int kSdfThreads = 320;
constexpr int kSdfStageBytes = 32768;                 // hi | lo of a 128(n) x 64(k) block
constexpr int kSdfStages = kSdfRingBytes / kSdfStageBytes;
constexpr float kSdfSA = 16.f;                        // activation scale
constexpr float kSdfSW = 64.f;                        // weight scale (robir_tc_pack_layer)
constexpr uint32_t kSdfIdesc = idesc_f16(128, 128);

struct SdfTcParams {
  const float* pts;      // [n][3]
  int n;
  float in_scale, sdf_scale, feat_scale;
  const uint8_t* img;    // 8 (+1 with features) layer images of 8 stages each, robir_tc_pack_layer(terms = 3)
  const float* bias;     // [8][256]
  const float* w8_sdf;   // [256]
  const float* b8;       // [257]
  float* sdf;            // [n]
  float* grad;           // [n][3] (jet mode) or null
  float* feat;           // [n][256] or null
  const int* n_active;   // optional device scalar (rows at or beyond it are skipped; the caller zero-fills)
};

// column i (0..63) of PE10(x) (value, jet = 0) or of its derivative with respect to x[jet - 1]
__device__ __forceinline__ float sdf_pe_elem(int i, const float (&x)[3], int jet) {
  if (i >= 63) return 0.f;
  if (i < 3) return jet == 0 ? x[i] : (jet - 1 == i ? 1.f : 0.f);
  const int l = (i - 3) / 6, rem = (i - 3) - 6 * l, d = rem >= 3 ? rem - 3 : rem;
  if (jet != 0 && jet - 1 != d) return 0.f;
  const float f = (float)(1 << l);
  const float arg = x[d] * f;
  if (jet == 0) return rem >= 3 ? cosf(arg) : sinf(arg);
  return rem >= 3 ? -f * sinf(arg) : f * cosf(arg);
}

template <bool JET, bool FEAT>
__global__ void __launch_bounds__(kSdfThreads, 1) sdf_tc_kernel(SdfTcParams p) {
  constexpr int kLayers = FEAT ? 9 : 8;
  constexpr int PTS = JET ? 32 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kSdfStages], empty_bar[kSdfStages], s_ready, a_ready[4], d_full[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[8 * 256];
  __shared__ __align__(16) float s_w8[256];
  __shared__ __align__(16) float s_b8f[256];
  __shared__ float s_part[128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kSdfStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&s_ready, 8);
    for (int k = 0; k < 4; ++k) mbar_init(&a_ready[k], 8);
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 8 * 256; i += kSdfThreads) s_bias[i] = p.bias[i];
  for (int i = tid; i < 256; i += kSdfThreads) { s_w8[i] = p.w8_sdf[i]; s_b8f[i] = p.b8[1 + i]; }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int n_act = p.n_active ? min(__ldg(p.n_active), p.n) : p.n;
  const int ntiles = (n_act + PTS - 1) / PTS;

  if (warp == 0) {
    // ===================================== weight producer =====================================
    int st = 0;
    uint32_t ph = 1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int layer = 0; layer < kLayers; ++layer) {
        const uint8_t* src = p.img + (size_t)layer * 8 * kSdfStageBytes;
        for (int j = 0; j < 8; ++j) {
          if (layer == 0 && (j & 5) != 0) continue;            // K = 64: only the K-block-0 stages (j = 0, 2)
          mbar_wait(&empty_bar[st], ph);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[st], kSdfStageBytes);
            bulk_g2s(ring + (size_t)st * kSdfStageBytes, src + (size_t)j * kSdfStageBytes, kSdfStageBytes, &full_bar[st]);
          }
          __syncwarp();
          if (++st == kSdfStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    int st = 0;
    uint32_t ph = 0, a_phase = 0, tile_it = 0;
    const uint64_t desc0 = smem_desc_sw128(ring);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
#pragma unroll 1
      for (int layer = 0; layer < kLayers; ++layer) {
        const uint32_t a_half = tmem_base + ((layer & 1) ? 256u : 0u);
        const uint32_t d_half = tmem_base + ((layer & 1) ? 0u : 256u);
        const uint32_t ready_par = (layer == 0 ? tile_it : a_phase) & 1;
        if (layer > 0) ++a_phase;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          constexpr int kNhOf[8] = {0, 0, 1, 1, 0, 0, 1, 1}, kKbOf[8] = {0, 1, 0, 1, 2, 3, 2, 3};
          const int nh = kNhOf[j], kb = kKbOf[j];
          if (layer == 0 && kb != 0) continue;
          if (nh == 0) {
            if (layer == 0) mbar_wait(&s_ready, ready_par); else mbar_wait(&a_ready[kb], ready_par);
          }
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t b_hi = desc0 + (uint64_t)((st * kSdfStageBytes) >> 4), b_lo = b_hi + (16384u >> 4);
            const uint32_t d_addr = d_half + 128u * nh;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              const uint32_t a_hi = a_half + 64u * kb + 32u * (s4 >> 1) + 8u * (s4 & 1), a_lo = a_hi + 16u;
              if (kb == 0 && s4 == 0) umma_ts<0>(d_addr, a_hi, b_hi + 2u * s4, kSdfIdesc);
              else umma_ts<1>(d_addr, a_hi, b_hi + 2u * s4, kSdfIdesc);
              umma_ts<1>(d_addr, a_lo, b_hi + 2u * s4, kSdfIdesc);
              umma_ts<1>(d_addr, a_hi, b_lo + 2u * s4, kSdfIdesc);
            }
            umma_commit(&empty_bar[st]);
            if (layer == 0) {
              if (j == 0) umma_commit(&d_full[0]);
              if (j == 2) umma_commit(&d_full[1]);
            } else {
              if (j == 5) umma_commit(&d_full[0]);
              if (j == 7) umma_commit(&d_full[1]);
            }
          }
          __syncwarp();
          if (++st == kSdfStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================================== epilogue warps =====================================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;                  // which 32 of the 64 columns of every K block this thread converts
    const int row = q * 32 + lane;
    const int jet = JET ? (row & 3) : 0;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float inv = 1.f / (kSdfSA * kSdfSW);
    uint32_t d_phase[2] = {0, 0};
    auto store_a = [&](uint32_t taddr, float (&v)[32]) {      // true-unit values -> scaled fp16 hi/lo A operand chunk
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) split_pack_f16_rz<false>(v[2 * i] * kSdfSA, v[2 * i + 1] * kSdfSA, hi[i], lo[i]);
      tmem_st16(taddr, hi);
      tmem_st16(taddr + 16u, lo);
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int pt = tile * PTS + (JET ? (row >> 2) : row);
      const bool valid = pt < n_act;
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = __ldg(p.pts + 3 * (size_t)pt + c) * p.in_scale;
      }
      // ---- first A operand: PE rows, K block 0 (columns [0, 64) of X)
      {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? sdf_pe_elem(32 * ch + i, x, jet) : 0.f;
        store_a(tmem_base + lane_addr + (uint32_t)(32 * ch), v);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_ready);
      }
      float dot = 0.f;
#pragma unroll 1
      for (int layer = 0; layer < kLayers; ++layer) {
        const uint32_t d_half = tmem_base + ((layer & 1) ? 0u : 256u);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&d_full[h], d_phase[h] & 1);
          ++d_phase[h];
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int kb = 2 * h + c;
            const int n0 = 64 * kb + 32 * ch;                  // first output column of this chunk
            const uint32_t taddr = d_half + lane_addr + (uint32_t)n0;
            uint32_t r[32];                                    // one chunk at a time: the softplus-jet epilogue needs
            tmem_ld32(taddr, r);                               // the registers (no spills) more than the overlap
            tmem_wait_ld();
            float v[32];
            if (layer < 8) {
              // Softplus(100) on the value rows, its derivative (of the VALUE row, same column) on the tangent rows
              const float* bp = s_bias + layer * 256 + n0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float a = __uint_as_float(r[i]) * inv;
                const float z = a + bp[i];                     // meaningful on value rows
                const float z0 = JET ? __shfl_sync(0xffffffffu, z, lane & ~3) : z;
                const float e = __expf(-100.f * fabsf(z0));
                if (jet == 0) {
                  v[i] = fmaxf(z0, 0.f) + __logf(1.f + e) * 0.01f;
                } else {
                  const float sg = z0 >= 0.f ? __fdividef(1.f, 1.f + e) : __fdividef(e, 1.f + e);
                  v[i] = sg * a;
                }
              }
              if (layer == 3 && n0 >= 192) {                   // cat([h (193 columns), PE (63 columns)]) (/ sqrt 2 in W4)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (n0 + i >= 193) v[i] = valid ? sdf_pe_elem(n0 + i - 193, x, jet) : 0.f;
              }
              if (layer == 7) {                                // layer 8, column 0: fp32 dot with the folded row
                const float4* wp = reinterpret_cast<const float4*>(s_w8 + n0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 w = wp[i];
                  dot = fmaf(w.x, v[4 * i], dot); dot = fmaf(w.y, v[4 * i + 1], dot);
                  dot = fmaf(w.z, v[4 * i + 2], dot); dot = fmaf(w.w, v[4 * i + 3], dot);
                }
              }
              if (layer < 7 || FEAT) {
                store_a(taddr, v);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[kb]);
              }
            } else if (FEAT) {
              // layer 8, feature columns 1..256 (value rows only)
              if (valid && jet == 0) {
                float* out = p.feat + (size_t)pt * 256 + n0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float4 o;
                  o.x = (__uint_as_float(r[4 * i]) * inv + s_b8f[n0 + 4 * i]) * p.feat_scale;
                  o.y = (__uint_as_float(r[4 * i + 1]) * inv + s_b8f[n0 + 4 * i + 1]) * p.feat_scale;
                  o.z = (__uint_as_float(r[4 * i + 2]) * inv + s_b8f[n0 + 4 * i + 2]) * p.feat_scale;
                  o.w = (__uint_as_float(r[4 * i + 3]) * inv + s_b8f[n0 + 4 * i + 3]) * p.feat_scale;
                  reinterpret_cast<float4*>(out)[i] = o;
                }
              }
            }
          }
        }
        if (layer == 7) {
          // the two column groups of a row combine their partial dots through shared memory (64-thread named barrier)
          if (ch == 1) s_part[row] = dot;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          if (ch == 0 && valid) {
            const float d = dot + s_part[row];
            if (jet == 0) p.sdf[pt] = (d + p.b8[0]) * p.sdf_scale;
            else if (p.grad) p.grad[3 * (size_t)pt + jet - 1] = d;   // d f / d p = d net / d x for in_scale * sdf_scale = 1
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // s_part is rewritten by the next tile
        }
      }
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

template <bool JET, bool FEAT>
static int launch_sdf_tc(const SdfTcParams& p, int grid, void* stream) {
  constexpr int smem = kSdfRingBytes + 1024;
  RB_CHECK_CUDA(cudaFuncSetAttribute(sdf_tc_kernel<JET, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  sdf_tc_kernel<JET, FEAT><<<grid, kSdfThreads, smem, (cudaStream_t)stream>>>(p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" {

// SDF network on the tensor-core engine: sdf [n] (always), grad [n][3] (optional -> jet mode), feat [n][256] (optional).
// img: 8 (9 with features) layer images packed with robir_tc_pack_layer(Wt_l, ldw = 256, N = 256, K = 64 | 256,
// transpose = 1, n_halves = 2, terms = 3) from the folded, transposed weights [K][256] (layer 4 pre-multiplied by
// 1/sqrt(2)); bias [8][256].  Rows at or beyond *n_active (if given) are not evaluated: the caller zero-fills.
int robir_sdf_tc(const SdfTcParams* p, int sm_count, void* stream) {
  if (p->n == 0) return 0;
  const bool jet = p->grad != nullptr, feat = p->feat != nullptr;
  const int pts = jet ? 32 : 128;
  const int tiles = (p->n + pts - 1) / pts;
  const int grid = tiles < sm_count ? tiles : sm_count;
  if (jet) return feat ? launch_sdf_tc<true, true>(*p, grid, stream) : launch_sdf_tc<true, false>(*p, grid, stream);
  return feat ? launch_sdf_tc<false, true>(*p, grid, stream) : launch_sdf_tc<false, false>(*p, grid, stream);
}

}  // extern "C"
