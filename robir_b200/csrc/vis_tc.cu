// Tensor-core (tcgen05 / TMEM) engine for the fused visibility MLP, forward and input-gradient backward.
//
// One CTA per SM (320 threads), persistent over 128-row tiles of the (point, direction) pair list.  Warp roles:
//   warp 0    : weight producer -- 1-D bulk async copies (UBLKCP) of pre-swizzled fp16 weight images from L2 into a
//               shared-memory ring (192 KB; a stage = B_hi [| B_lo] of a 128(n) x 64(k) block, SWIZZLE_128B K-major);
//   warp 1    : MMA issuer -- one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16);
//   warps 2-9 : epilogue -- two warps per TMEM lane quarter, each thread owns one row and 64 of the 128 columns of an
//               N-half: tcgen05.ld the fp32 accumulators, bias + ReLU (fwd) / ReLU-mask (bwd), split into fp16 hi/lo and
//               tcgen05.st them back IN PLACE as the next layer's A operand.
// Activations never touch shared or global memory: the A operand of every layer lives in TMEM (two 256-column halves
// X / Y used ping-pong as "A of this layer" / "D of this layer"), so all of shared memory feeds weights.
//
// Number format.  TERMS = 3 (fp32 parity, engine "tc"): every product is evaluated as hi*hi + lo*hi + hi*lo on fp16
// operands with fp32 accumulation.  fp16 hi+lo carries 22 mantissa bits (bf16 hi+lo: 16) as long as the values sit in
// fp16's normal range, which power-of-two scales guarantee: activations are carried as 16 x, weights as 64 w, the
// backward chain starts from the unit gradient 256 * wd (the per-row scalar dL/dlogit multiplies the result in fp32 at
// the very end), so nothing under- or overflows for |activation| < 4e3, |weight| < 1e3.  Measured against fp64 on the
// CPU emulation (tools/vis_numerics_study.py): outputs 1e-7, input gradient 3e-7 relative -- the fp32 path's own error
// -- where the former bf16 split had 2e-6 / 2e-3 (ReLU units flipping sign).  TERMS = 1 (engine "tc1", fast mode):
// single-pass fp16, 1/3 of the MMAs, half the weight stream; error ~1e-4 (TF32-class), NOT the parity mode.
//
// Schedule.  Within a layer the eight (n-half, k-block) stages run K-half-major: (n0,k0) (n0,k1) (n1,k0) (n1,k1)
// (n0,k2) (n0,k3) (n1,k2) (n1,k3).  The first four only need K-half 0 of the A operand, i.e. the epilogue of N-half 0
// of the layer before, which itself finished two stages before that layer's end; the epilogue of N-half 1 runs under
// them.  Across tiles: X's K-half 0 is dead after the last layer's 4th stage (x_free), so the next tile's first A
// operand (gather of the layer-0 tables) is written while the last layer still computes, and the next tile's first
// MMAs start as soon as the epilogue has drained the D half they overwrite (d_free).  The tensor pipe never waits for
// an epilogue that is not at least one K-half of MMAs ahead.
//
// TMEM layout of an A half (256 K-values of 128 rows in 256 columns): 32-column chunk c' holds K in [32c', 32c'+32):
//   columns [32c', 32c'+16) = packed fp16 hi (2 K-values per column), [32c'+16, 32c'+32) = packed fp16 lo (TERMS = 3),
// which is exactly where the epilogue finds the fp32 accumulators of the 32 output features it converts.
#include "common.cuh"
#include "tc_common.cuh"

namespace robir {
using namespace tc;

constexpr int kTcRingBytes = 196608;
constexpr int kTcThreads = 320;
constexpr float kSA = 16.f;     // activation scale
constexpr float kSW = 64.f;     // weight scale
constexpr float kSG = 256.f;    // unit-gradient scale of the backward chain
constexpr uint32_t kIdescN128 = idesc_f16(128, 128);
constexpr uint32_t kIdescN64 = idesc_f16(128, 64);

__host__ __device__ constexpr int tc_stage_bytes(int terms, int rows) { return rows * 128 * (terms == 3 ? 2 : 1); }
// streaming order of the (n-half, k-block) stages of a 256x256 layer
__host__ __device__ constexpr int tc_stage_nh(int j) { return (j >> 1) & 1; }
__host__ __device__ constexpr int tc_stage_kb(int j) { return (j & 1) | ((j >> 2) << 1); }
__host__ __device__ constexpr int tc_stage_of(int nh, int kb) { return ((kb >> 1) << 2) | (nh << 1) | (kb & 1); }

// ------------------------------------------------------------------------------------------------------------------
// weight image of one layer: stages in streaming order; a stage = [hi: rows x 64 k][lo: rows x 64 k (TERMS = 3)] fp16,
// SWIZZLE_128B (16-byte chunk c of row r stored at chunk c ^ (r & 7)).  B[n][k] = scale * (transpose ? W[k][n] : W[n][k]);
// rows >= N or k >= K are zero.  n_halves = 2: 128-row stages, 8 of them; n_halves = 1: 64-row stages, 4 of them.
// ------------------------------------------------------------------------------------------------------------------
__global__ void pack_tc_image_kernel(const float* __restrict__ W, int ldw, int N, int K, int transpose, int n_halves,
                                     int terms, float scale, uint8_t* __restrict__ img) {
  const int rows = n_halves == 2 ? 128 : 64;
  // one thread per 16-byte chunk: 8 consecutive k of one row, hi and lo
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n_halves * 4 * rows * 8;
  if (idx >= total) return;
  const int chunk = idx & 7;
  const int r = (idx >> 3) % rows;
  const int rest = (idx >> 3) / rows;
  const int kb = rest & 3, nh = rest >> 2;
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
      x[e] = v * scale;
    }
    split_pack_f16(x[0], x[1], hi[j], lo[j]);
  }
  const int j = n_halves == 2 ? tc_stage_of(nh, kb) : kb;
  uint8_t* stage = img + (size_t)j * tc_stage_bytes(terms, rows);
  const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(stage + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (terms == 3) *reinterpret_cast<uint4*>(stage + rows * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct TcParams {
  // pair list (128-row tiles)
  const int* rowA; const int* rowB; const int* n_tiles;
  // forward
  const float* tabA; const float* tabB;
  const uint8_t* img;      // fwd: 3 layers x 8 stages; bwd: 3 layers x 8 stages + 4 64-row stages (W0d)
  const float* bias;       // [3][256] (fwd)
  const float* wd;         // [256]
  const float* bd;         // [1]
  float* vis;              // [rows]
  uint32_t* mask;          // [tiles][4 layers][8 words][128 rows] (a warp stores one 128-byte line); bit (31 - i) of
                           // word w of layer l = (pre-activation of h_{l+1}[32 w + i] is not negative)
  // backward
  const float* g_vis; const float* dirs; float* g_dirs;
  // self-test: plain GEMM D = A . W^T through the same machinery
  const float* test_A; float* test_D;
  // dynamic tile scheduler: zero-initialised counter; CTA b starts on tile b and then draws gridDim.x + counter++
  int* tile_counter;
  // optional per-CTA stall accounting (robir_tc_debug_buffer): [grid][8] clock counts, see tools/tc_stalls.py
  unsigned long long* dbg;
};

__device__ __forceinline__ long long tc_clock() { return clock64(); }
// wait on an mbarrier; with PROF the waiting time is added to `acc` (tools/tc_microbench.py)
#define TC_WAIT(acc, bar, par)                  \
  do {                                          \
    if (PROF) {                                 \
      const long long _t = tc_clock();          \
      mbar_wait(bar, par);                      \
      acc += tc_clock() - _t;                   \
    } else {                                    \
      mbar_wait(bar, par);                      \
    }                                           \
  } while (0)

template <int MODE, int TERMS, bool PROF>  // MODE 0 = forward, 1 = backward, 2 = self-test (one 256x256 layer)
__global__ void __launch_bounds__(kTcThreads, 1) vis_tc_kernel(TcParams p) {
  constexpr int SB = tc_stage_bytes(TERMS, 128);        // bytes of a 128-row stage
  constexpr int SB64 = tc_stage_bytes(TERMS, 64);
  constexpr int NST = kTcRingBytes / SB;                // 6 (TERMS = 3) or 12
  constexpr int kLayers = (MODE == 0) ? 3 : (MODE == 1 ? 4 : 1);
  constexpr int kXLayer = 2;                            // the layer during which X's K-half 0 dies (fwd: last, bwd: L2)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // A-operand hand-offs per 64-wide K block: s_ready[kb] = K block kb of a tile's FIRST operand (written one tile
  // ahead), a_ready[kb] = K block kb produced by a layer epilogue of the current tile (separate barriers: in the
  // backward the next tile's first operand is ready before this tile's last hand-off).
  __shared__ uint64_t full_bar[NST], empty_bar[NST], s_ready[4], a_ready[4], d_full[2], d_free[2], x_free;
  // tile scheduler: the epilogue's leader thread draws the next tile one tile ahead and publishes it to all roles
  __shared__ uint64_t t_ready[2];
  __shared__ int s_tile[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[3 * 256];
  __shared__ __align__(16) float s_wd[256];
  __shared__ float s_part[2][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int k = 0; k < 4; ++k) { mbar_init(&a_ready[k], 8); mbar_init(&s_ready[k], 8); }   // one arrival per epilogue warp
    mbar_init(&d_full[0], 1); mbar_init(&d_full[1], 1);
    mbar_init(&d_free[0], 8); mbar_init(&d_free[1], 8);
    mbar_init(&x_free, 1);
    mbar_init(&t_ready[0], 1); mbar_init(&t_ready[1], 1);
    fence_barrier_init();
  }
  if (MODE == 0) for (int i = tid; i < 3 * 256; i += kTcThreads) s_bias[i] = p.bias[i] * kSA;
  if (MODE <= 1) for (int i = tid; i < 256; i += kTcThreads) s_wd[i] = p.wd[i];
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int ntiles = (MODE >= 2) ? 1 : *p.n_tiles;

  if (warp == 0) {
    // ===================================== weight producer =====================================
    int st = 0;
    uint32_t ph = 1;                                     // "slot is free" parity (fresh barriers pass)
    long long w_empty = 0;
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile >= 0 && tile < ntiles; ++tile_it) {
      for (int layer = 0; layer < kLayers; ++layer) {
        const bool last64 = (MODE == 1 && layer == 3);
        const int nst = last64 ? 4 : 8;
        const uint32_t bytes = last64 ? SB64 : SB;
        const uint8_t* src = p.img + (size_t)layer * 8 * SB;
        for (int j = 0; j < nst; ++j) {
          TC_WAIT(w_empty, &empty_bar[st], ph);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[st], bytes);
            bulk_g2s(ring + (size_t)st * SB, src + (size_t)j * bytes, bytes, &full_bar[st]);
          }
          __syncwarp();
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
      if (MODE >= 2) break;
      mbar_wait(&t_ready[(tile_it + 1) & 1], (tile_it >> 1) & 1);          // next tile of this CTA (or -1)
      tile = s_tile[(tile_it + 1) & 1];
    }
    if (PROF && p.dbg && lane == 0) p.dbg[blockIdx.x * 8 + 7] = (unsigned long long)w_empty;
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // The whole warp runs the loop converged and one elected lane issues (elect.sync): tcgen05.mma / commit are
    // warp-uniform instructions, and inside a divergent `if (lane == 0)` ptxas wraps every one of them in an
    // ELECT / BRA.U.ANY retry loop (~75 issue cycles per MMA).  The stage loop is fully unrolled so that the (n-half,
    // k-block) of a stage, the barriers it waits on and the ones it commits are compile-time constants.
    int st = 0;
    uint32_t ph = 0, a_phase = 0, tile_it = 0;
    long long w_full = 0, w_ready = 0, w_dfree = 0;
    const long long t_begin = PROF ? tc_clock() : 0;
    const uint64_t desc0 = smem_desc_sw128(ring);
    for (int tile = blockIdx.x; tile >= 0 && tile < ntiles; ++tile_it) {
#pragma unroll 1
      for (int layer = 0; layer < kLayers; ++layer) {
        uint64_t* ready = layer == 0 ? s_ready : a_ready;
        const uint32_t ready_par = (layer == 0 ? tile_it : a_phase) & 1;
        if (layer > 0) ++a_phase;
        const uint32_t a_half = tmem_base + ((layer & 1) ? 256u : 0u);
        const uint32_t d_half = tmem_base + ((layer & 1) ? 0u : 256u);
        if (MODE == 1 && layer == 3) {
          // ---- backward tail: N = 64 (W0d^T), four 64-row stages; the 64 dPE columns go to X[128, 192)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            TC_WAIT(w_ready, &ready[j], ready_par);
            TC_WAIT(w_full, &full_bar[st], ph);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint64_t b_hi = desc0 + (uint64_t)((st * SB) >> 4), b_lo = b_hi + (8192u >> 4);
              const uint32_t d_addr = tmem_base + 128u;
#pragma unroll
              for (int s4 = 0; s4 < 4; ++s4) {
                const uint32_t a_hi = a_half + 64u * j + 32u * (s4 >> 1) + 8u * (s4 & 1), a_lo = a_hi + 16u;
                if (j == 0 && s4 == 0) umma_ts<0>(d_addr, a_hi, b_hi + 2u * s4, kIdescN64);
                else umma_ts<1>(d_addr, a_hi, b_hi + 2u * s4, kIdescN64);
                if (TERMS == 3) {
                  umma_ts<1>(d_addr, a_lo, b_hi + 2u * s4, kIdescN64);
                  umma_ts<1>(d_addr, a_hi, b_lo + 2u * s4, kIdescN64);
                }
              }
              umma_commit(&empty_bar[st]);
              if (j == 3) umma_commit(&d_full[0]);
            }
            __syncwarp();
            if (++st == NST) { st = 0; ph ^= 1; }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          constexpr int kNhOf[8] = {0, 0, 1, 1, 0, 0, 1, 1}, kKbOf[8] = {0, 1, 0, 1, 2, 3, 2, 3};
          const int nh = kNhOf[j], kb = kKbOf[j];
          if (nh == 0) TC_WAIT(w_ready, &ready[kb], ready_par);                   // K block kb of A is in TMEM
          if (MODE == 0 && kb == 0) { if (layer == 0) TC_WAIT(w_dfree, &d_free[nh], (tile_it & 1) ^ 1); }  // D drained
          TC_WAIT(w_full, &full_bar[st], ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t b_hi = desc0 + (uint64_t)((st * SB) >> 4), b_lo = b_hi + (16384u >> 4);
            const uint32_t d_addr = d_half + 128u * nh;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              const uint32_t a_hi = a_half + 64u * kb + 32u * (s4 >> 1) + 8u * (s4 & 1), a_lo = a_hi + 16u;
              if (kb == 0 && s4 == 0) umma_ts<0>(d_addr, a_hi, b_hi + 2u * s4, kIdescN128);
              else umma_ts<1>(d_addr, a_hi, b_hi + 2u * s4, kIdescN128);
              if (TERMS == 3) {
                umma_ts<1>(d_addr, a_lo, b_hi + 2u * s4, kIdescN128);
                umma_ts<1>(d_addr, a_hi, b_lo + 2u * s4, kIdescN128);
              }
            }
            umma_commit(&empty_bar[st]);
            if (j == 5) umma_commit(&d_full[0]);
            if (j == 7) umma_commit(&d_full[1]);
            if (MODE <= 1 && j == 3) { if (layer == kXLayer) umma_commit(&x_free); }
          }
          __syncwarp();
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
      if (MODE >= 2) break;
      mbar_wait(&t_ready[(tile_it + 1) & 1], (tile_it >> 1) & 1);
      tile = s_tile[(tile_it + 1) & 1];
    }
    if (PROF && p.dbg && lane == 0) {
      unsigned long long* d = p.dbg + blockIdx.x * 8;
      d[0] = (unsigned long long)(tc_clock() - t_begin); d[1] = (unsigned long long)w_full;
      d[2] = (unsigned long long)w_ready; d[3] = (unsigned long long)w_dfree;
    }
  } else {
    // ===================================== epilogue warps =====================================
    // Thread (row, ch): row = its TMEM lane, ch = which 32 of the 64 columns of every K block [64 kb, 64 kb + 64) it
    // converts: one 32-column chunk per hand-off, four hand-offs per layer.
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float inv_sw = 1.f / kSW;
    uint32_t d_phase[2] = {0, 0}, x_phase = 0;
    long long w_dfull = 0, w_xfree = 0;
    const long long t_begin = PROF ? tc_clock() : 0;
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
    // v (scaled; forward: pre-activations, relu applied by the converts; backward: masked gradients) -> A operand chunk
    auto store_a = [&](uint32_t taddr, float (&v)[32]) {
      uint32_t hi[16], lo[16];
      if (TERMS == 3) {
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pack_f16_rz<MODE == 0>(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
        tmem_st16(taddr, hi);
        tmem_st16(taddr + 16u, lo);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) hi[i] = pack_f16_act<MODE == 0>(v[2 * i], v[2 * i + 1]);
        tmem_st16(taddr, hi);
      }
    };
    // ---- first A operand of a tile, K block kb: this thread's 32 columns [64 kb + 32 ch, +32) of X
    float4 ga[2][4], gb[2][4];
    auto gather = [&](int col, int slot, int a, int b) {      // 16 features of tabA[a] and tabB[b] from column col
      const float4* pa = reinterpret_cast<const float4*>(p.tabA + (size_t)a * 256 + col);
      const float4* pb = reinterpret_cast<const float4*>(p.tabB + (size_t)(b >= 0 ? b : 0) * 256 + col);
#pragma unroll
      for (int i = 0; i < 4; ++i) { ga[slot][i] = __ldg(pa + i); gb[slot][i] = __ldg(pb + i); }
    };
    uint32_t mw0[4];                                           // backward: this thread's 4 words of mask slot 3
    auto stage0_prefetch = [&](int t, int kb, int a, int b) {  // requests what stage0_block(t, kb) consumes
      if (MODE == 0) { gather(64 * kb + 32 * ch, 0, a, b); gather(64 * kb + 32 * ch + 16, 1, a, b); }
      if (MODE == 1 && kb == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mw0[i] = __ldg(p.mask + ((size_t)t * 32 + 3 * 8 + 2 * i + ch) * 128 + row);
      }
    };
    // converts K block kb; with `more` the loads of block kb + 1 are requested while this one converts
    auto stage0_block = [&](int t, int kb, int a, int b, bool more) {
      const bool ok = b >= 0;
      const int col = 64 * kb + 32 * ch;
      float x[32];
      if (MODE == 0) {
        uint32_t m = 0;
#pragma unroll
        for (int s = 0; s < 2; ++s) {                          // two 16-feature gathers per 32-column chunk
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 va = ga[s][i], vb = gb[s][i];
            const float2 sa2 = make_float2(kSA, kSA);
            const float2 t0 = ffma2(make_float2(va.x, va.y), sa2, make_float2(vb.x * kSA, vb.y * kSA));
            const float2 t1 = ffma2(make_float2(va.z, va.w), sa2, make_float2(vb.z * kSA, vb.w * kSA));
            m = __funnelshift_l(__float_as_uint(t0.x), m, 1);
            m = __funnelshift_l(__float_as_uint(t0.y), m, 1);
            m = __funnelshift_l(__float_as_uint(t1.x), m, 1);
            m = __funnelshift_l(__float_as_uint(t1.y), m, 1);
            x[16 * s + 4 * i] = t0.x; x[16 * s + 4 * i + 1] = t0.y;
            x[16 * s + 4 * i + 2] = t1.x; x[16 * s + 4 * i + 3] = t1.y;
          }
          if (more) gather(col + 64 + 16 * s, s, a, b);
        }
        if (p.mask && ok) p.mask[((size_t)t * 32 + (col >> 5)) * 128 + row] = ~m;
      } else if (MODE == 1) {
        const uint32_t mword = ok ? mw0[kb] : 0u;
        const float4* wp = reinterpret_cast<const float4*>(s_wd + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = wp[i];
          x[4 * i] = ((int)(mword << (4 * i)) < 0) ? kSG * w.x : 0.f;
          x[4 * i + 1] = ((int)(mword << (4 * i + 1)) < 0) ? kSG * w.y : 0.f;
          x[4 * i + 2] = ((int)(mword << (4 * i + 2)) < 0) ? kSG * w.z : 0.f;
          x[4 * i + 3] = ((int)(mword << (4 * i + 3)) < 0) ? kSG * w.w : 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = p.test_A[(size_t)row * 256 + col + i] * kSA;
      }
      store_a(tmem_base + lane_addr + (uint32_t)col, x);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_ready[kb]);
    };

    if (blockIdx.x < ntiles) {
      if (MODE <= 1) load_row(blockIdx.x, a_idx, b_idx, g0);
      stage0_prefetch(blockIdx.x, 0, a_idx, b_idx);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) stage0_block(blockIdx.x, kb, a_idx, b_idx, kb < 3);
    }
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile >= 0 && tile < ntiles; ++tile_it) {
      const int q0 = tile * 128;
      const bool valid = b_idx >= 0;
      int tile_nxt = -1;
      bool has_next = false;
      uint32_t* mtile = (MODE == 0 && p.mask && valid) ? p.mask + (size_t)tile * (32 * 128) + row : nullptr;
      int a_nxt = 0, b_nxt = -1;
      float g_nxt = 0.f;
      float logit = 0.f;
#pragma unroll 1
      for (int layer = 0; layer < kLayers; ++layer) {
        const uint32_t d_half = tmem_base + ((layer & 1) ? 0u : 256u);
        const bool last64 = (MODE == 1 && layer == 3);
        const bool fwd_last = (MODE == 0 && layer == 2);
        if (MODE <= 1 && layer == kXLayer - 1) {
          // draw the next tile (one thread), publish it to the producer / issuer / epilogue threads of this CTA
          if (warp == 2 && lane == 0) {
            int t = (int)gridDim.x + atomicAdd(p.tile_counter, 1);
            s_tile[(tile_it + 1) & 1] = t < ntiles ? t : -1;
            __threadfence_block();
            mbar_arrive(&t_ready[(tile_it + 1) & 1]);
          }
          mbar_wait(&t_ready[(tile_it + 1) & 1], (tile_it >> 1) & 1);
          tile_nxt = s_tile[(tile_it + 1) & 1];
          has_next = tile_nxt >= 0;
          if (has_next) load_row(tile_nxt, a_nxt, b_nxt, g_nxt);
        }
        if (has_next && layer == kXLayer) {
          // next tile's first A operand, K blocks 0 and 1: X[0, 128) is dead once this layer's 4th stage has completed
          stage0_prefetch(tile_nxt, 0, a_nxt, b_nxt);
          TC_WAIT(w_xfree, &x_free, x_phase & 1);
          ++x_phase;
          tc_fence_after();
          stage0_block(tile_nxt, 0, a_nxt, b_nxt, true);
          stage0_block(tile_nxt, 1, a_nxt, b_nxt, false);
        }
        // backward: the 4 mask words of this layer's epilogue (slot 2 - layer), requested before the wait
        uint32_t mw[4] = {0u, 0u, 0u, 0u};
        if (MODE == 1 && layer < 3 && valid) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            mw[i] = __ldg(p.mask + ((size_t)tile * 32 + (2 - layer) * 8 + 2 * i + ch) * 128 + row);
        }
        if (last64) {
          // ---- backward tail (epilogue group 0 only): dPE[0..63] in X[128, 192) -> d dir via the PE jacobian.
          // Every epilogue thread observes the phase (a waiter may never fall two phases behind an mbarrier, or the
          // parity test of its next wait aliases); group 1 must not overwrite X[160, 192) before group 0 has read it.
          TC_WAIT(w_dfull, &d_full[0], d_phase[0] & 1);
          ++d_phase[0];
          tc_fence_after();
          uint32_t r0[32], r1[32];
          if (ch == 0) {
            tmem_ld32(tmem_base + 128u + lane_addr, r0);
            tmem_ld32(tmem_base + 160u + lane_addr, r1);
            tmem_wait_ld();
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          if (ch == 0 && valid) {
            float dpe[64];
#pragma unroll
            for (int i = 0; i < 32; ++i) { dpe[i] = __uint_as_float(r0[i]); dpe[32 + i] = __uint_as_float(r1[i]); }
            const float gs = g0 * (inv_sw / kSG);            // accumulators carry kSG * kSW * (unit gradient)
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
              atomicAdd(p.g_dirs + 3 * b_idx + i, g * gs);
            }
          }
          continue;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // D columns [128 h, 128 h + 128) are complete: both of this thread's chunks of the half are requested at once
          TC_WAIT(w_dfull, &d_full[h], d_phase[h] & 1);
          ++d_phase[h];
          tc_fence_after();
          // software pipeline over the two chunks of the half: the second chunk's accumulators are requested right
          // after the first chunk's have arrived and travel while the first chunk converts (TMEM reads are 64 B/clk:
          // requesting both up front doubles the time to the first hand-off)
          uint32_t rr[2][32];
          tmem_ld32(d_half + lane_addr + (uint32_t)(128 * h + 32 * ch), rr[0]);
          tmem_wait_ld();
          tmem_ld32(d_half + lane_addr + (uint32_t)(128 * h + 64 + 32 * ch), rr[1]);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int kb = 2 * h + c;
            const int n0 = 64 * kb + 32 * ch;                  // first output feature of this chunk
            const uint32_t taddr = d_half + lane_addr + (uint32_t)n0;
            if (c == 1) tmem_wait_ld();
            uint32_t (&r)[32] = rr[c];
            float x[32];
            if (MODE == 0) {
              const float2* bp = reinterpret_cast<const float2*>(s_bias + layer * 256 + n0);
              uint32_t m = 0;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 v = ffma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])),
                                       make_float2(inv_sw, inv_sw), bp[i]);
                m = __funnelshift_l(__float_as_uint(v.x), m, 1);
                m = __funnelshift_l(__float_as_uint(v.y), m, 1);
                x[2 * i] = v.x; x[2 * i + 1] = v.y;            // relu happens in the converts (store_a) / below
              }
              if (mtile != nullptr) mtile[((layer + 1) * 8 + (n0 >> 5)) * 128] = ~m;
            } else if (MODE == 1) {
              const uint32_t mword = mw[kb];
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = ((int)(mword << i) < 0) ? __uint_as_float(r[i]) * inv_sw : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                p.test_D[(size_t)row * 256 + n0 + i] = __uint_as_float(r[i]) * (inv_sw / kSA);
            }
            if (fwd_last) {
              const float4* wp = reinterpret_cast<const float4*>(s_wd + n0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 w = wp[i];
                logit = fmaf(w.x, fmaxf(x[4 * i], 0.f), logit); logit = fmaf(w.y, fmaxf(x[4 * i + 1], 0.f), logit);
                logit = fmaf(w.z, fmaxf(x[4 * i + 2], 0.f), logit); logit = fmaf(w.w, fmaxf(x[4 * i + 3], 0.f), logit);
              }
            } else if (MODE <= 1) {
              store_a(taddr, x);
              tmem_wait_st();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[kb]);
            }
          }
          if (fwd_last) {                                      // this D half may be overwritten by the next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_free[h]);
          }
        }
      }
      if (MODE == 0) {
        // the two column groups of a row combine their partial logits through shared memory (64-thread named barrier)
        if (ch == 1) s_part[tile_it & 1][row] = logit;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        if (ch == 0)
          p.vis[q0 + row] = valid ? 1.f / (1.f + expf(-((logit + s_part[tile_it & 1][row]) * (1.f / kSA) + p.bd[0]))) : 0.f;
      }
      if (has_next) {                                          // X[128, 256): dead since the last layer completed
        stage0_prefetch(tile_nxt, 2, a_nxt, b_nxt);
        stage0_block(tile_nxt, 2, a_nxt, b_nxt, true);
        stage0_block(tile_nxt, 3, a_nxt, b_nxt, false);
      }
      a_idx = a_nxt; b_idx = b_nxt; g0 = g_nxt;
      if (MODE >= 2) break;
      tile = tile_nxt;
    }
    if (PROF && p.dbg && warp == 2 && lane == 0) {
      unsigned long long* d = p.dbg + blockIdx.x * 8;
      d[4] = (unsigned long long)(tc_clock() - t_begin); d[5] = (unsigned long long)w_dfull;
      d[6] = (unsigned long long)w_xfree;
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

static constexpr int kTcSmem = kTcRingBytes + 1024;

static unsigned long long* g_tc_dbg = nullptr;
static int* g_tile_counters = nullptr;       // pool of tile-scheduler counters (one per launch in flight, round robin)
static unsigned g_tile_counter_next = 0;
constexpr int kTileCounters = 256;

template <int MODE, int TERMS>
static int launch_tc(TcParams p, int grid, void* stream) {
  p.dbg = g_tc_dbg;
  if (g_tile_counters == nullptr) RB_CHECK_CUDA(cudaMalloc(&g_tile_counters, kTileCounters * sizeof(int)));
  p.tile_counter = g_tile_counters + (g_tile_counter_next++ % kTileCounters);
  RB_CHECK_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int), (cudaStream_t)stream));
  if (g_tc_dbg != nullptr) {          // stall accounting build of the same kernel (clock reads around every wait)
    RB_CHECK_CUDA(cudaFuncSetAttribute(vis_tc_kernel<MODE, TERMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kTcSmem));
    vis_tc_kernel<MODE, TERMS, true><<<grid, kTcThreads, kTcSmem, (cudaStream_t)stream>>>(p);
  } else {
    RB_CHECK_CUDA(cudaFuncSetAttribute(vis_tc_kernel<MODE, TERMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kTcSmem));
    vis_tc_kernel<MODE, TERMS, false><<<grid, kTcThreads, kTcSmem, (cudaStream_t)stream>>>(p);
  }
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" {

// Bytes of the weight image of n256 256x256 layers followed by n64 64-row layers.  terms = 3 (parity) or 1 (fast).
int robir_tc_image_bytes(int n_layers256, int n_layers64, int terms) {
  return n_layers256 * 8 * tc_stage_bytes(terms, 128) + n_layers64 * 4 * tc_stage_bytes(terms, 64);
}

// Packs one layer's weights into the tensor-core image.  transpose = 0: B[n][k] = W[n][k] (forward, W = torch weight
// [N][K]); transpose = 1: B[n][k] = W[k][n] (backward through the same layer).  n_halves = 2 for 256 rows, 1 for <= 64.
int robir_tc_pack_layer(const float* W, int ldw, int N, int K, int transpose, int n_halves, int terms, void* img,
                        void* stream) {
  RB_REQUIRE(n_halves == 1 || n_halves == 2, "tc_pack_layer: n_halves must be 1 or 2");
  RB_REQUIRE(terms == 1 || terms == 3, "tc_pack_layer: terms must be 1 (fast) or 3 (fp32 parity)");
  RB_REQUIRE(K <= 256 && N <= (n_halves == 2 ? 256 : 64), "tc_pack_layer: shape exceeds the image");
  const int total = n_halves * 4 * (n_halves == 2 ? 128 : 64) * 8;
  pack_tc_image_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, transpose, n_halves, terms,
                                                                               kSW, (uint8_t*)img);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Forward over a 128-row-tile pair list.  img: 3 layers (robir_tc_pack_layer, transpose=0, same terms) back to back.
int robir_vis_tc_fwd(const float* tabA, const float* tabB, const int* rowA, const int* rowB, const int* n_tiles,
                     int max_tiles, const void* img, const float* bias3x256, const float* wd, const float* bd,
                     float* vis, uint32_t* mask, int terms, int sm_count, void* stream) {
  if (max_tiles == 0) return 0;
  RB_REQUIRE(terms == 1 || terms == 3, "vis_tc_fwd: terms must be 1 or 3");
  TcParams p = {};
  p.rowA = rowA; p.rowB = rowB; p.n_tiles = n_tiles; p.tabA = tabA; p.tabB = tabB; p.img = (const uint8_t*)img;
  p.bias = bias3x256; p.wd = wd; p.bd = bd; p.vis = vis; p.mask = mask;
  const int grid = max_tiles < sm_count ? max_tiles : sm_count;
  return terms == 3 ? launch_tc<0, 3>(p, grid, stream) : launch_tc<0, 1>(p, grid, stream);
}

// Backward (input gradient w.r.t. the direction).  img: W3^T, W2^T, W1^T (transpose=1, 2 halves each) then the
// 64-row W0d image (robir_tc_pack_layer(W0 + 63, ldw=126, N=63, K=256, transpose=1, n_halves=1)).
int robir_vis_tc_bwd(const int* rowB, const int* n_tiles, int max_tiles, const void* img, const float* wd,
                     const float* vis, const float* g_vis, const uint32_t* mask, const float* dirs, float* g_dirs,
                     int terms, int sm_count, void* stream) {
  if (max_tiles == 0) return 0;
  RB_REQUIRE(terms == 1 || terms == 3, "vis_tc_bwd: terms must be 1 or 3");
  TcParams p = {};
  p.rowB = rowB; p.n_tiles = n_tiles; p.img = (const uint8_t*)img; p.wd = wd; p.vis = const_cast<float*>(vis);
  p.g_vis = g_vis; p.mask = const_cast<uint32_t*>(mask); p.dirs = dirs; p.g_dirs = g_dirs;
  const int grid = max_tiles < sm_count ? max_tiles : sm_count;
  return terms == 3 ? launch_tc<1, 3>(p, grid, stream) : launch_tc<1, 1>(p, grid, stream);
}

// Diagnostics: device buffer of [sm_count][8] 64-bit clock counts that every following vis_tc launch overwrites
// (issuer: total, wait weights, wait A operand, wait D drained; epilogue warp 2: total, wait accumulators, wait X free;
// producer: wait ring slot); NULL switches it off.  tools/tc_stalls.py prints the breakdown.
int robir_tc_debug_buffer(void* buf) {
  g_tc_dbg = (unsigned long long*)buf;
  return 0;
}

// Self-test of the GEMM machinery: D[128][256] = A[128][256] . W[256][256]^T (img packed with transpose=0, one layer).
int robir_tc_selftest(const float* A, const void* img, float* D, int terms, void* stream) {
  RB_REQUIRE(terms == 1 || terms == 3, "tc_selftest: terms must be 1 or 3");
  TcParams p = {};
  p.img = (const uint8_t*)img; p.test_A = A; p.test_D = D;
  return terms == 3 ? launch_tc<2, 3>(p, 1, stream) : launch_tc<2, 1>(p, 1, stream);
}

}  // extern "C"
