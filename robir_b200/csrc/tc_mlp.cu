// Tensor-core (tcgen05 / TMEM) layer engine for the 512-wide material / indirect-illumination networks (SURVEY.md rows
// a6 / a7): one launch per Linear layer, one CTA per (128-row tile, 128-column block) of the output.
//
//   forward  : Y = act(A . W^T + b)                      A = previous layer's activations
//   backward : G_prev = (G . W) * act'(saved A_prev)     (input-gradient chain; weight gradients: wgrad_kernel, mlp.cu)
//
// Both operands are bf16 hi/lo images in global memory (L2-resident), K-major SWIZZLE_128B tiles that are exactly the
// shared-memory layout tcgen05.mma reads, so a pipeline stage is two 1-D bulk copies (UBLKCP): the A k-block of this
// row tile (128 rows x 64 k: hi 16 KB | lo 16 KB) and the W k-block of this column block (same shape).  Every product
// is hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (fp32 parity, like vis_tc.cu).  The epilogue (one thread per
// row) applies bias / activation (or the activation derivative), writes the fp32 rows that the backward and the weight
// gradients need, and emits the hi/lo image of its 128 output columns = two k-blocks of the next layer's A operand.
// Warp roles: warp 0 producer, warp 1 MMA issuer (elect.sync lane), warps 2-9 epilogue (TMEM lane quarter warp % 4,
// column half (warp - 2) / 4).
#include "common.cuh"
#include "tc_common.cuh"

namespace robir {
using namespace tc;

constexpr int kTlStages = 3;
constexpr int kTlStageBytes = 65536;     // A k-block (32 KB) + W k-block (32 KB)
constexpr int kTlBlockBytes = 32768;     // one 128 x 64 hi|lo k-block
constexpr int kTlThreads = 320;       // producer, MMA issuer, 8 epilogue warps (two per scheduler: 64 columns each)
constexpr uint32_t kTlIdesc = idesc_bf16(128, 128);
constexpr int kTlSmem = kTlStages * kTlStageBytes + 1024 + 8 * 4096;      // ring + alignment + epilogue staging

struct TcLayerParams {
  const uint8_t* a_img;      // [row_tiles][nkb][32 KB]
  const uint8_t* w_img;      // [col_blocks][nkb][32 KB]
  const float* bias;         // [>= 128 * col_blocks] (zero padded) or null
  int n, N, nkb;             // rows, valid output columns, k-blocks of the contraction
  int mode;                  // 0 forward, 1 backward
  int act;                   // forward: activation of this layer; backward: activation whose derivative is applied
  const float* ref;          // backward: saved post-activation values of the previous layer [n][ld_ref] (null: none)
  int ld_ref;
  float* out;                // fp32 rows [n][ld_out], columns < N written (null: skip)
  int ld_out;
  uint8_t* out_img;          // [row_tiles][nkb_out][32 KB] image of the output (null: skip)
  int nkb_out;
  const int* n_active;       // see MlpParams (mlp.cu)
  int seg;
  int no_fill;               // 1: outputs of inactive row tiles are left unwritten (every consumer skips those rows too)
};

__device__ __forceinline__ float tl_act(float x, int act) {
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_LEAKY02) return x > 0.f ? x : 0.2f * x;
  // softplus(beta = 100) as max(x, 0) + log(1 + e^(-100 |x|)) / 100: two MUFU ops, absolute error < 1e-8 (the log's
  // argument is in [1, 2]); agrees with common.cuh's thresholded log1p(exp) form to that bound
  if (act == ACT_SOFTPLUS100) return fmaxf(x, 0.f) + 0.01f * __logf(1.f + __expf(-100.f * fabsf(x)));
  return x;
}
__device__ __forceinline__ float tl_dact(float y, int act) {
  if (act == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == ACT_LEAKY02) return y > 0.f ? 1.f : 0.2f;
  // softplus(beta = 100): y = log(1 + e^(100 x)) / 100  =>  sigmoid(100 x) = 1 - e^(-100 y)
  if (act == ACT_SOFTPLUS100) return 1.f - __expf(-100.f * y);          // absolute error <= 1 ulp of 1
  return 1.f;
}

// Epilogue stores go through a per-warp 4 KB staging buffer in shared memory: a thread owns one accumulator row, so a
// direct store instruction would touch 32 different 128-byte lines with 16 bytes each (32 LSU wavefronts, half-written
// sectors); after the transposition each instruction writes whole lines (fp32 rows: 4 lines per instruction) or whole
// 64-byte halves of image rows (8 per instruction).
constexpr int kTlStageWarpBytes = 4096;

// hi / lo words of 32 consecutive columns of this warp's 32 rows (thread = row q * 32 + lane) -> image k-block;
// col0 % 32 == 0; warp-collective
__device__ __forceinline__ void tl_store_image(uint8_t* img_tile, int nkb_out, int r, int lane, int col0,
                                               const float (&x)[32], uint8_t* stage) {
  const int kb = col0 >> 6;
  if (kb >= nkb_out) return;
  const int c16_0 = (col0 & 63) >> 3;                 // first 16-byte chunk (8 bf16) of this 32-column group
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_pack(x[8 * j + 2 * e], x[8 * j + 2 * e + 1], hi[e], lo[e]);
    const int off = lane * 64 + ((j ^ sw) << 4);
    *reinterpret_cast<uint4*>(stage + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(stage + 2048 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  __syncwarp();
  uint8_t* blk = img_tile + (size_t)kb * kTlBlockBytes + (r - lane) * 128;      // this warp's first row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = i * 8 + (lane >> 2), j = lane & 3;
    const int src = rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4);
    const int dst = rr * 128 + (((c16_0 + j) ^ (rr & 7)) << 4);                  // (r - lane) % 8 == 0
    *reinterpret_cast<uint4*>(blk + dst) = *reinterpret_cast<const uint4*>(stage + src);
    *reinterpret_cast<uint4*>(blk + 16384 + dst) = *reinterpret_cast<const uint4*>(stage + 2048 + src);
  }
  __syncwarp();
}

// 32 fp32 columns of this warp's 32 rows -> out rows (whole 128-byte lines per instruction); warp-collective
__device__ __forceinline__ void tl_store_rows(float* dst_w0 /* row of lane 0, column col0 */, int ld_out, int rows_valid,
                                              int lane, const float (&x)[32], uint8_t* stage) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3), pc = lane & 7;
    const float4 v = *reinterpret_cast<const float4*>(stage + rr * 128 + ((pc ^ (rr & 7)) << 4));
    if (rr < rows_valid) *reinterpret_cast<float4*>(dst_w0 + (size_t)rr * ld_out + pc * 4) = v;
  }
  __syncwarp();
}

// activation / activation derivative of 32 values with the activation as a compile-time constant: straight-line code
// whose 32 independent MUFU / FMA chains interleave (a per-element switch on p.act serialises them)
template <int ACT>
__device__ __forceinline__ void tl_act32(float (&x)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = tl_act(x[i], ACT);
}
template <int ACT>
__device__ __forceinline__ void tl_dact32(float (&d)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) d[i] = tl_dact(d[i], ACT);
}

// one 32-column chunk of the epilogue: accumulator columns at taddr (this warp's lane quarter) -> bias / activation
// (forward) or activation derivative (backward) -> fp32 rows and / or the hi|lo image of the next layer's A operand
__device__ __forceinline__ void tl_epilogue_chunk(const TcLayerParams& p, uint32_t taddr, int row, int r, int lane,
                                                  int col0, uint8_t* img_tile, uint8_t* stage) {
  uint32_t acc[32];
  tmem_ld32(taddr, acc);
  float x[32];
  const bool full = col0 + 32 <= p.N;               // whole chunk inside the valid columns (the common case)
  if (p.mode == 0) {
    if (p.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);   // zero padded to 128
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = 0.f;
    }
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] += __uint_as_float(acc[i]);
    switch (p.act) {
      case ACT_RELU: tl_act32<ACT_RELU>(x); break;
      case ACT_LEAKY02: tl_act32<ACT_LEAKY02>(x); break;
      case ACT_SOFTPLUS100: tl_act32<ACT_SOFTPLUS100>(x); break;
      default: break;
    }
    if (!full) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i >= p.N) x[i] = 0.f;
    }
  } else {
    const bool have_ref = p.ref != nullptr && row < p.n;
    if (have_ref && full && (p.ld_ref & 3) == 0) {
      const float4* rp = reinterpret_cast<const float4*>(p.ref + (size_t)row * p.ld_ref + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(rp + i);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
    } else if (have_ref) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = col0 + i < p.N ? __ldg(p.ref + (size_t)row * p.ld_ref + col0 + i) : 0.f;
    }
    if (have_ref) {
      switch (p.act) {
        case ACT_RELU: tl_dact32<ACT_RELU>(x); break;
        case ACT_LEAKY02: tl_dact32<ACT_LEAKY02>(x); break;
        case ACT_SOFTPLUS100: tl_dact32<ACT_SOFTPLUS100>(x); break;
        default: tl_dact32<0>(x); break;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = 1.f;
    }
    tmem_wait_ld();
    const bool row_ok = row < p.n;
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (row_ok && (full || col0 + i < p.N)) ? __uint_as_float(acc[i]) * x[i] : 0.f;
  }
  if (p.out != nullptr) {
    if (full && (p.ld_out & 3) == 0) {
      tl_store_rows(p.out + (size_t)(row - lane) * p.ld_out + col0, p.ld_out, p.n - (row - lane), lane, x, stage);
    } else if (row < p.n) {
      float* dst = p.out + (size_t)row * p.ld_out + col0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) dst[i] = x[i];
    }
  }
  if (img_tile != nullptr) tl_store_image(img_tile, p.nkb_out, r, lane, col0, x, stage);
}

__global__ void __launch_bounds__(kTlThreads, 1) tc_layer_kernel(TcLayerParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kTlStages], empty_bar[kTlStages], d_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, cb = blockIdx.y;        // row tile, 128-column block
  const int row0 = tile * 128, col_base = cb * 128;

  // ---- inactive row tile (hit rays are compacted to the front of every segment): zero outputs, no arithmetic
  {
    const int sg = p.seg > 0 ? p.seg : (p.n > 0 ? p.n : 1);
    const int n_act = p.n_active ? min(__ldg(p.n_active), sg) : sg;
    const int o = row0 % sg;
    if (o >= n_act && o + 128 <= sg) {
      if (p.no_fill) return;
      if (p.out != nullptr)
        for (int i = tid; i < 128 * 128; i += kTlThreads) {
          const int r = row0 + (i >> 7), c = col_base + (i & 127);
          if (r < p.n && c < p.N) p.out[(size_t)r * p.ld_out + c] = 0.f;
        }
      if (p.out_img != nullptr)
        for (int kb = 2 * cb; kb < 2 * cb + 2 && kb < p.nkb_out; ++kb) {
          uint4* dst = reinterpret_cast<uint4*>(p.out_img + ((size_t)tile * p.nkb_out + kb) * kTlBlockBytes);
          for (int i = tid; i < kTlBlockBytes / 16; i += kTlThreads) dst[i] = make_uint4(0, 0, 0, 0);
        }
      return;
    }
  }

  if (tid == 0) {
    for (int s = 0; s < kTlStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================================== producer =====================================
    const uint8_t* a_src = p.a_img + (size_t)tile * p.nkb * kTlBlockBytes;
    const uint8_t* w_src = p.w_img + (size_t)cb * p.nkb * kTlBlockBytes;
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int st = kb % kTlStages;
      mbar_wait(&empty_bar[st], ((kb / kTlStages) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[st], kTlStageBytes);
        bulk_g2s(ring + (size_t)st * kTlStageBytes, a_src + (size_t)kb * kTlBlockBytes, kTlBlockBytes, &full_bar[st]);
        bulk_g2s(ring + (size_t)st * kTlStageBytes + kTlBlockBytes, w_src + (size_t)kb * kTlBlockBytes, kTlBlockBytes,
                 &full_bar[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int st = kb % kTlStages;
      mbar_wait(&full_bar[st], (kb / kTlStages) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint8_t* sa = ring + (size_t)st * kTlStageBytes;
        const uint8_t* sb = sa + kTlBlockBytes;
        const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + 16384);
        const uint64_t b_hi = smem_desc_sw128(sb), b_lo = smem_desc_sw128(sb + 16384);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (kb == 0 && q == 0) umma_ss<0>(tmem_base, a_hi + 2u * q, b_hi + 2u * q, kTlIdesc);
          else umma_ss<1>(tmem_base, a_hi + 2u * q, b_hi + 2u * q, kTlIdesc);
          umma_ss<1>(tmem_base, a_lo + 2u * q, b_hi + 2u * q, kTlIdesc);
          umma_ss<1>(tmem_base, a_hi + 2u * q, b_lo + 2u * q, kTlIdesc);
        }
        umma_commit(&empty_bar[st]);
        if (kb == p.nkb - 1) umma_commit(&d_full);
      }
      __syncwarp();
    }
  } else {
    // ===================================== epilogue =====================================
    const int q = warp & 3;
    const int r = q * 32 + lane, row = row0 + r;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* img_tile = p.out_img ? p.out_img + (size_t)tile * p.nkb_out * kTlBlockBytes : nullptr;
    mbar_wait(&d_full, 0);
    tc_fence_after();
    const int half = (warp - 2) >> 2;
#pragma unroll 1
    uint8_t* stage = ring + kTlStages * kTlStageBytes + (warp - 2) * kTlStageWarpBytes;
    for (int c = 2 * half; c < 2 * half + 2; ++c)
      tl_epilogue_chunk(p, tmem_base + lane_addr + 32u * c, row, r, lane, col_base + 32 * c, img_tile, stage);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------------------------
// the same layer for large row counts (the CESR chains: n_hit x 128 rows): persistent CTAs, 128 x 256 output tiles
// (N = 256 MMAs: half the A-operand traffic of 128 x 128 tiles), two TMEM accumulators so that the epilogue of tile i
// runs under the MMAs of tile i + 1.  Stage = A k-block (hi 16 KB | lo 16 KB) + W k-blocks of the two 128-column
// blocks (hi 32 KB | lo 32 KB) = 96 KB, two stages.  Tiles are scheduled column-tile fastest: the CTAs that share a row
// tile's A image read it back to back (L2), the 1 MB weight image stays L2-resident.  A tile whose second column block
// does not exist (N <= 128 mod 256) runs N = 128 MMAs.  Inactive row tiles (see tc_layer_kernel) are zero-filled by
// the epilogue warps and skipped by the producer / MMA warps.
constexpr int kTbStages = 2;
constexpr int kTbStageBytes = 98304;
constexpr int kTbSmem = kTbStages * kTbStageBytes + 1024 + 8 * 4096;
constexpr uint32_t kTbIdesc256 = idesc_bf16(128, 256);

__device__ __forceinline__ bool tb_tile_active(const TcLayerParams& p, int row0) {
  const int sg = p.seg > 0 ? p.seg : (p.n > 0 ? p.n : 1);
  const int n_act = p.n_active ? min(__ldg(p.n_active), sg) : sg;
  const int o = row0 % sg;
  return !(o >= n_act && o + 128 <= sg);
}

__global__ void __launch_bounds__(kTlThreads, 1) tc_layer_big_kernel(TcLayerParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kTbStages], empty_bar[kTbStages], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col_blocks = (p.N + 127) >> 7, col_tiles = (col_blocks + 1) >> 1, row_tiles = (p.n + 127) >> 7;
  const int n_tiles = row_tiles * col_tiles;

  if (tid == 0) {
    for (int s = 0; s < kTbStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_free[b], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================================== producer =====================================
    int it = 0;                                             // k-block counter over this CTA's active tiles
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int tile = t / col_tiles, ct = t % col_tiles;
      if (!tb_tile_active(p, tile * 128)) continue;
      const int ncb = min(2, col_blocks - 2 * ct);
      const uint8_t* a_src = p.a_img + (size_t)tile * p.nkb * kTlBlockBytes;
      const uint8_t* w_src = p.w_img + (size_t)(2 * ct) * p.nkb * kTlBlockBytes;
      for (int kb = 0; kb < p.nkb; ++kb, ++it) {
        const int st = it % kTbStages;
        mbar_wait(&empty_bar[st], ((it / kTbStages) & 1) ^ 1);
        if (elect_one_sync()) {
          uint8_t* sa = ring + (size_t)st * kTbStageBytes;
          mbar_arrive_expect_tx(&full_bar[st], kTlBlockBytes * (1 + ncb));
          bulk_g2s(sa, a_src + (size_t)kb * kTlBlockBytes, kTlBlockBytes, &full_bar[st]);
          for (int j = 0; j < ncb; ++j) {
            const uint8_t* wb = w_src + ((size_t)j * p.nkb + kb) * kTlBlockBytes;
            bulk_g2s(sa + 32768 + j * 16384, wb, 16384, &full_bar[st]);                    // hi rows 128 j ..
            bulk_g2s(sa + 65536 + j * 16384, wb + 16384, 16384, &full_bar[st]);            // lo rows 128 j ..
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    int it = 0, j = 0;                                      // k-block / active-tile counters
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int tile = t / col_tiles, ct = t % col_tiles;
      if (!tb_tile_active(p, tile * 128)) continue;
      const int ncb = min(2, col_blocks - 2 * ct);
      const uint32_t idesc = ncb == 2 ? kTbIdesc256 : kTlIdesc;
      const int buf = j & 1;
      mbar_wait(&acc_free[buf], ((j >> 1) & 1) ^ 1);        // the epilogue of the tile before last has drained it
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 256u * buf;
      for (int kb = 0; kb < p.nkb; ++kb, ++it) {
        const int st = it % kTbStages;
        mbar_wait(&full_bar[st], (it / kTbStages) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint8_t* sa = ring + (size_t)st * kTbStageBytes;
          const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + 16384);
          const uint64_t b_hi = smem_desc_sw128(sa + 32768), b_lo = smem_desc_sw128(sa + 65536);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (kb == 0 && q == 0) umma_ss<0>(d_tmem, a_hi + 2u * q, b_hi + 2u * q, idesc);
            else umma_ss<1>(d_tmem, a_hi + 2u * q, b_hi + 2u * q, idesc);
            umma_ss<1>(d_tmem, a_lo + 2u * q, b_hi + 2u * q, idesc);
            umma_ss<1>(d_tmem, a_hi + 2u * q, b_lo + 2u * q, idesc);
          }
          umma_commit(&empty_bar[st]);
          if (kb == p.nkb - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
      ++j;
    }
  } else {
    // ===================================== epilogue =====================================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* stage = ring + kTbStages * kTbStageBytes + (warp - 2) * kTlStageWarpBytes;
    int j = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int tile = t / col_tiles, ct = t % col_tiles;
      const int row0 = tile * 128, row = row0 + r;
      const int ncb = min(2, col_blocks - 2 * ct);
      uint8_t* img_tile = p.out_img ? p.out_img + (size_t)tile * p.nkb_out * kTlBlockBytes : nullptr;
      if (!tb_tile_active(p, row0)) {
        // zero outputs of an inactive row tile: this warp's 32 rows x its 128-column block, coalesced
        if (half < ncb && !p.no_fill) {
          const int col_base = (2 * ct + half) * 128;
          if (p.out != nullptr) {
            const bool vec = (p.ld_out & 3) == 0 && col_base + 128 <= p.N;
            for (int rr = 0; rr < 32; ++rr) {
              const int orow = row0 + q * 32 + rr;
              if (orow >= p.n) break;
              float* d = p.out + (size_t)orow * p.ld_out + col_base;
              if (vec) {
                reinterpret_cast<float4*>(d)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
              } else {
                for (int c = lane; c < 128 && col_base + c < p.N; c += 32) d[c] = 0.f;
              }
            }
          }
          if (img_tile != nullptr)
            for (int kb = 2 * (2 * ct + half); kb < 2 * (2 * ct + half) + 2 && kb < p.nkb_out; ++kb) {
              uint8_t* blk = img_tile + (size_t)kb * kTlBlockBytes + q * 4096;       // rows 32 q .. of the hi half
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                *reinterpret_cast<uint4*>(blk + (i * 32 + lane) * 16) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(blk + 16384 + (i * 32 + lane) * 16) = make_uint4(0, 0, 0, 0);
              }
            }
        }
        continue;
      }
      const int buf = j & 1;
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      if (half < ncb) {
        const int col_base = (2 * ct + half) * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c)
          tl_epilogue_chunk(p, tmem_base + lane_addr + 256u * buf + 128u * half + 32u * c, row, r, lane,
                            col_base + 32 * c, img_tile, stage);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free[buf]);
      ++j;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// images
// ------------------------------------------------------------------------------------------------------------------
// weights: B[nn][kk] = transpose ? W[kk][nn] : W[nn][kk] (W row-major, leading dimension ldw), nn < N, kk < K, zero
// padded to [col_blocks * 128][nkb * 64]; layout [col_block][kb][hi 16 KB | lo 16 KB], SWIZZLE_128B rows of 64 k
__global__ void tl_pack_weight_kernel(const float* __restrict__ W, int ldw, int N, int K, int transpose, int col_blocks,
                                      int nkb, uint8_t* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)col_blocks * nkb * 128 * 8;
  if (idx >= total) return;
  const int chunk = idx & 7, r = (idx >> 3) & 127;
  const int blk = (int)(idx >> 10), kb = blk % nkb, cbk = blk / nkb;
  const int nn = cbk * 128 + r;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = kb * 64 + chunk * 8 + 2 * q + e;
      float v = 0.f;
      if (nn < N && kk < K) v = transpose ? W[(size_t)kk * ldw + nn] : W[(size_t)nn * ldw + kk];
      x[e] = v;
    }
    split_pack(x[0], x[1], hi[q], lo[q]);
  }
  uint8_t* b = img + (size_t)blk * kTlBlockBytes;
  const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(b + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(b + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// activations: fp32 rows X [n][ldx] (columns < K valid, optionally multiplied by act'(ref)) -> image
// [row_tiles][nkb][hi | lo]; rows >= n and columns >= K are zero
__global__ void tl_pack_rows_kernel(const float* __restrict__ X, int ldx, int n, int K, const float* __restrict__ ref,
                                    int ld_ref, int act, int nkb, uint8_t* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int tiles = (n + 127) / 128;
  const long long total = (long long)tiles * nkb * 128 * 8;
  if (idx >= total) return;
  const int chunk = idx & 7, r = (idx >> 3) & 127;
  const int blk = (int)(idx >> 10), kb = blk % nkb, tile = blk / nkb;
  const int row = tile * 128 + r;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = kb * 64 + chunk * 8 + 2 * q + e;
      float v = 0.f;
      if (row < n && kk < K) {
        v = X[(size_t)row * ldx + kk];
        if (ref != nullptr) v *= tl_dact(ref[(size_t)row * ld_ref + kk], act);
      }
      x[e] = v;
    }
    split_pack(x[0], x[1], hi[q], lo[q]);
  }
  uint8_t* b = img + (size_t)blk * kTlBlockBytes;
  const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(b + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(b + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradients on the tensor cores: dW [N][K] = G^T A (contraction over the n rows), db = column sums of G
// ------------------------------------------------------------------------------------------------------------------
// Both operands become images of their TRANSPOSES (image row = column of the fp32 matrix, k = its row), so that the
// contraction index is the K-major one and the GEMM below is the same two-bulk-copies-per-stage SS pipeline as
// tc_layer_kernel.  One CTA transposes a 64-row x 128-column fp32 tile through shared memory: coalesced loads, full
// 128-byte image rows out.  With colsum != null it also emits the tile's column sums (db partials, summed in a fixed
// order by tl_wgrad_reduce_kernel).
__global__ void __launch_bounds__(256) tl_pack_rows_t_kernel(const float* __restrict__ X, int ldx, int n, int C, int nkb,
                                                             uint8_t* __restrict__ img, float* __restrict__ colsum,
                                                             int ld_colsum, const int* __restrict__ n_active) {
  __shared__ float t[64][129];
  const int kb = blockIdx.x, ct = blockIdx.y, tid = threadIdx.x;
  const int r0 = kb * 64, c0 = ct * 128;
  if (n_active != nullptr) {                  // rows >= *n_active are zero by contract: their k-blocks are never read
    n = min(n, __ldg(n_active));
    if (r0 >= n) return;
  }
  const bool vec = (ldx & 3) == 0 && c0 + 128 <= C && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (vec) {
    for (int i = tid; i < 64 * 32; i += 256) {
      const int r = i >> 5, c4 = (i & 31) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < n) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)(r0 + r) * ldx + c0 + c4));
      t[r][c4] = v.x; t[r][c4 + 1] = v.y; t[r][c4 + 2] = v.z; t[r][c4 + 3] = v.w;
    }
  } else {
    for (int i = tid; i < 64 * 128; i += 256) {
      const int r = i >> 7, c = i & 127;
      t[r][c] = (r0 + r < n && c0 + c < C) ? __ldg(X + (size_t)(r0 + r) * ldx + c0 + c) : 0.f;
    }
  }
  __syncthreads();
  if (colsum != nullptr && tid < 128) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < 64; ++r) s += t[r][tid];
    colsum[(size_t)kb * ld_colsum + c0 + tid] = s;
  }
  uint8_t* blk = img + ((size_t)ct * nkb + kb) * kTlBlockBytes;
  for (int i = tid; i < 128 * 8; i += 256) {
    const int chunk = i & 7, c = i >> 3;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_pack(t[chunk * 8 + 2 * e][c], t[chunk * 8 + 2 * e + 1][c], hi[e], lo[e]);
    const int off = c * 128 + ((chunk ^ (c & 7)) << 4);
    *reinterpret_cast<uint4*>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(blk + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct TlWgradParams {
  const uint8_t* g_img;      // image of G^T: [row_tiles][nkb][32 KB]   (row = output feature)
  const uint8_t* a_img;      // image of A^T: [col_blocks][nkb][32 KB]  (row = input feature)
  int nkb, kb_per_split;     // k-blocks (64 rows of G / A each) in total / per CTA
  int row_tiles, col_blocks;
  float* partial;            // [splits][row_tiles * 128][col_blocks * 128]
  const int* n_active;       // device row count (rows beyond it are zero and skipped), or null
};

// grid (row_tiles * col_blocks, splits): one 128 x 128 tile of dW over k-blocks [split * kb_per_split, ...)
__global__ void __launch_bounds__(192, 1) tl_wgrad_kernel(TlWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kTlStages], empty_bar[kTlStages], d_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / p.col_blocks, cb = blockIdx.x % p.col_blocks;
  int nkb_eff = p.nkb, kps = p.kb_per_split;
  if (p.n_active != nullptr) {                // the active k-blocks are re-divided over the launched splits
    nkb_eff = min(p.nkb, (max(__ldg(p.n_active), 0) + 63) >> 6);
    kps = (nkb_eff + (int)gridDim.y - 1) / (int)gridDim.y;
  }
  const int kb0 = blockIdx.y * kps, nk = min(nkb_eff - kb0, kps);     // <= 0: this split has no rows, partial = 0

  if (tid == 0) {
    for (int s = 0; s < kTlStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&d_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    const uint8_t* g_src = p.g_img + ((size_t)tile * p.nkb + kb0) * kTlBlockBytes;
    const uint8_t* a_src = p.a_img + ((size_t)cb * p.nkb + kb0) * kTlBlockBytes;
    for (int kb = 0; kb < nk; ++kb) {
      const int st = kb % kTlStages;
      mbar_wait(&empty_bar[st], ((kb / kTlStages) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[st], kTlStageBytes);
        bulk_g2s(ring + (size_t)st * kTlStageBytes, g_src + (size_t)kb * kTlBlockBytes, kTlBlockBytes, &full_bar[st]);
        bulk_g2s(ring + (size_t)st * kTlStageBytes + kTlBlockBytes, a_src + (size_t)kb * kTlBlockBytes, kTlBlockBytes,
                 &full_bar[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    for (int kb = 0; kb < nk; ++kb) {
      const int st = kb % kTlStages;
      mbar_wait(&full_bar[st], (kb / kTlStages) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint8_t* sa = ring + (size_t)st * kTlStageBytes;
        const uint8_t* sb = sa + kTlBlockBytes;
        const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + 16384);
        const uint64_t b_hi = smem_desc_sw128(sb), b_lo = smem_desc_sw128(sb + 16384);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (kb == 0 && q == 0) umma_ss<0>(tmem_base, a_hi + 2u * q, b_hi + 2u * q, kTlIdesc);
          else umma_ss<1>(tmem_base, a_hi + 2u * q, b_hi + 2u * q, kTlIdesc);
          umma_ss<1>(tmem_base, a_lo + 2u * q, b_hi + 2u * q, kTlIdesc);
          umma_ss<1>(tmem_base, a_hi + 2u * q, b_lo + 2u * q, kTlIdesc);
        }
        umma_commit(&empty_bar[st]);
        if (kb == nk - 1) umma_commit(&d_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int ld = p.col_blocks * 128;
    float* dst = p.partial + ((size_t)blockIdx.y * p.row_tiles * 128 + tile * 128 + r) * ld + cb * 128;
    if (nk > 0) {
      mbar_wait(&d_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t acc[32];
      if (nk > 0) {
        tmem_ld32(tmem_base + lane_addr + 32u * c, acc);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(dst + 32 * c)[i] =
            make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                        __uint_as_float(acc[4 * i + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradients straight from the layer engine's own images (no transposing pack): the K-major SWIZZLE_128B block of
// 128 rows x 64 features IS the MN-major SWIZZLE_128B atom sequence the MMA wants when the contraction runs over the
// rows -- 64 contiguous features (128 bytes) per row, eight rows per 1024-byte swizzle atom, the XOR on the row index.
// So dW[n][k] = sum_r G[r][n] A[r][k] takes both operands MN-major (instruction descriptor bits 15 / 16) from the image
// of G (the backward chain's input image of this layer) and the image of A (the forward chain's input image).
// Stage = 64 rows: G features [128 tn, +128) as two 64-feature pieces (hi 8 KB each | lo 8 KB each) + the same for A
// features [128 tk, +128) = 64 KB; LBO = 8 KB (next 64 features), SBO = 1 KB (next 8 rows), +2 KB per 16-row k-step.
// db rides along: CTAs with tk == 0 multiply G by a tile of ones (N = 16) into 16 more TMEM columns.
struct TlWgradMnParams {
  const uint8_t* g_img;      // [row_tiles][nkb_g][32 KB] image of G (rows x N)
  const uint8_t* a_img;      // [row_tiles][nkb_a][32 KB] image of A (rows x K)
  int nkb_g, nkb_a, row_tiles;
  int n;                     // rows
  int tiles_n, tiles_k;      // 128-feature tiles of N / K
  float* partial;            // [splits][tiles_n * 128][tiles_k * 128]
  float* db_part;            // [splits][tiles_n * 128] or null
  const int* n_active;
};

__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N, int a_mn, int b_mn) {
  return idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(const void* base, uint32_t lbo_bytes) {
  const uint64_t addr = (uint64_t)((smem_u32(base) & 0x3FFFF) >> 4);
  return addr | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(192, 1) tl_wgrad_mn_kernel(TlWgradMnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kTlStages], empty_bar[kTlStages], d_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tn = blockIdx.x / p.tiles_k, tk = blockIdx.x % p.tiles_k;
  const bool with_db = p.db_part != nullptr && tk == 0;
  uint8_t* ones = ring + kTlStages * kTlStageBytes;                  // 2 KB of bf16 1.0 (16 rows x 128 B)

  int rows = p.n;
  if (p.n_active != nullptr) rows = min(rows, max(__ldg(p.n_active), 0));
  const int tiles_eff = min(p.row_tiles, (rows + 127) >> 7);
  const int per = (tiles_eff + (int)gridDim.y - 1) / (int)gridDim.y;
  const int t0 = blockIdx.y * per, t1 = min(tiles_eff, t0 + per);
  const int nsteps = max(t1 - t0, 0) * 2;                            // 64-row steps
  // feature pieces that exist in the images (the others stay unwritten in shared memory: they only feed output rows /
  // columns beyond N / K, which are not stored)
  const int g_pieces = min(2, p.nkb_g - 2 * tn), a_pieces = min(2, p.nkb_a - 2 * tk);

  if (tid == 0) {
    for (int s = 0; s < kTlStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&d_full, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 2048 / 4; i += 192) reinterpret_cast<uint32_t*>(ones)[i] = 0x3f803f80u;
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    for (int it = 0; it < nsteps; ++it) {
      const int st = it % kTlStages;
      mbar_wait(&empty_bar[st], ((it / kTlStages) & 1) ^ 1);
      if (elect_one_sync()) {
        const int tile = t0 + (it >> 1), half = it & 1;
        uint8_t* sg = ring + (size_t)st * kTlStageBytes;
        uint8_t* sa = sg + 32768;
        mbar_arrive_expect_tx(&full_bar[st], 16384u * (g_pieces + a_pieces));
        for (int c = 0; c < g_pieces; ++c) {
          const uint8_t* blk = p.g_img + ((size_t)tile * p.nkb_g + 2 * tn + c) * kTlBlockBytes + half * 8192;
          bulk_g2s(sg + c * 8192, blk, 8192, &full_bar[st]);                  // hi, rows 64 half ..
          bulk_g2s(sg + 16384 + c * 8192, blk + 16384, 8192, &full_bar[st]);  // lo
        }
        for (int c = 0; c < a_pieces; ++c) {
          const uint8_t* blk = p.a_img + ((size_t)tile * p.nkb_a + 2 * tk + c) * kTlBlockBytes + half * 8192;
          bulk_g2s(sa + c * 8192, blk, 8192, &full_bar[st]);
          bulk_g2s(sa + 16384 + c * 8192, blk + 16384, 8192, &full_bar[st]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t kIdesc = idesc_bf16_mn(128, 128, 1, 1);
    constexpr uint32_t kIdescOnes = idesc_bf16_mn(128, 16, 1, 0);
    const uint64_t ones_desc = smem_desc_sw128(ones);
    for (int it = 0; it < nsteps; ++it) {
      const int st = it % kTlStages;
      mbar_wait(&full_bar[st], (it / kTlStages) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint8_t* sg = ring + (size_t)st * kTlStageBytes;
        const uint8_t* sa = sg + 32768;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                       // 16 rows per k-step = 2 KB
          const uint64_t g_hi = smem_desc_sw128_mn(sg + q * 2048, 8192), g_lo = smem_desc_sw128_mn(sg + 16384 + q * 2048, 8192);
          const uint64_t a_hi = smem_desc_sw128_mn(sa + q * 2048, 8192), a_lo = smem_desc_sw128_mn(sa + 16384 + q * 2048, 8192);
          if (it == 0 && q == 0) umma_ss<0>(tmem_base, g_hi, a_hi, kIdesc);
          else umma_ss<1>(tmem_base, g_hi, a_hi, kIdesc);
          umma_ss<1>(tmem_base, g_lo, a_hi, kIdesc);
          umma_ss<1>(tmem_base, g_hi, a_lo, kIdesc);
          if (with_db) {
            if (it == 0 && q == 0) umma_ss<0>(tmem_base + 128u, g_hi, ones_desc, kIdescOnes);
            else umma_ss<1>(tmem_base + 128u, g_hi, ones_desc, kIdescOnes);
            umma_ss<1>(tmem_base + 128u, g_lo, ones_desc, kIdescOnes);
          }
        }
        umma_commit(&empty_bar[st]);
        if (it == nsteps - 1) umma_commit(&d_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int ld = p.tiles_k * 128;
    float* dst = p.partial + ((size_t)blockIdx.y * p.tiles_n * 128 + tn * 128 + r) * ld + tk * 128;
    if (nsteps > 0) {
      mbar_wait(&d_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t acc[32];
      if (nsteps > 0) {
        tmem_ld32(tmem_base + lane_addr + 32u * c, acc);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(dst + 32 * c)[i] =
            make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                        __uint_as_float(acc[4 * i + 3]));
    }
    if (with_db) {
      uint32_t acc[32];
      float v = 0.f;
      if (nsteps > 0) {
        tmem_ld32(tmem_base + lane_addr + 128u, acc);          // 16 valid columns (all equal: every column of the ones tile)
        tmem_wait_ld();
        v = __uint_as_float(acc[0]);
      }
      p.db_part[(size_t)blockIdx.y * p.tiles_n * 128 + tn * 128 + r] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// dW[nn][kk] = sum over splits (fixed order).  The last blocks of the grid sum the db partials instead: 8 columns x 32
// k-block lanes per block, lane j takes k-blocks j, j + 32, ..., then a fixed-order tree over the lanes.
__global__ void __launch_bounds__(256) tl_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int rows_pad,
                                                              int ld, int N, int K, float* __restrict__ dW,
                                                              const float* __restrict__ colsum, int nkb, int ld_colsum,
                                                              float* __restrict__ db, int dw_blocks,
                                                              const int* __restrict__ n_active) {
  if ((int)blockIdx.x < dw_blocks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * K) return;
    const int nn = (int)(idx / K), kk = (int)(idx % K);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)sp * rows_pad + nn) * ld + kk];
    dW[idx] = s;
    return;
  }
  __shared__ float red[32][9];
  if (n_active != nullptr) nkb = min(nkb, (max(__ldg(n_active), 0) + 63) >> 6);
  const int c = threadIdx.x & 7, j = threadIdx.x >> 3;
  const int nn = ((int)blockIdx.x - dw_blocks) * 8 + c;
  float s = 0.f;
  if (nn < N)
    for (int kb = j; kb < nkb; kb += 32) s += colsum[(size_t)kb * ld_colsum + nn];
  red[j][c] = s;
  __syncthreads();
  for (int w = 16; w >= 1; w >>= 1) {
    if (j < w) red[j][c] += red[j + w][c];
    __syncthreads();
  }
  if (j == 0 && nn < N) db[nn] = red[0][c];
}

}  // namespace robir

using namespace robir;

extern "C" {

int robir_tl_block_bytes(void) { return kTlBlockBytes; }

// W [N][K] (ldw) -> weight image [ceil(N'/128)][nkb][32 KB]; transpose = 1 packs W^T (N' = K rows, contraction over N)
int robir_tl_pack_weight(const float* W, int ldw, int N, int K, int transpose, int col_blocks, int nkb, void* img,
                         void* stream) {
  RB_REQUIRE(col_blocks * 128 >= N && nkb * 64 >= K, "tl_pack_weight: image smaller than the matrix");
  const long long total = (long long)col_blocks * nkb * 128 * 8;
  tl_pack_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, transpose,
                                                                                           col_blocks, nkb, (uint8_t*)img);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// fp32 rows -> activation image (optionally times act'(ref): the last layer's activation derivative of a backward chain)
int robir_tl_pack_rows(const float* X, int ldx, int n, int K, const float* ref, int ld_ref, int act, int nkb, void* img,
                       void* stream) {
  if (n == 0) return 0;
  RB_REQUIRE(nkb * 64 >= K, "tl_pack_rows: image smaller than the rows");
  const long long total = (long long)((n + 127) / 128) * nkb * 128 * 8;
  tl_pack_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, n, K, ref, ld_ref, act, nkb,
                                                                                        (uint8_t*)img);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_tl_layer(const TcLayerParams* p, void* stream) {
  if (p->n == 0) return 0;
  RB_REQUIRE(p->nkb >= 1 && p->nkb <= 8 && p->N >= 1, "tl_layer: 1..8 k-blocks (K <= 512)");
  RB_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTlSmem));
  dim3 grid((p->n + 127) / 128, (p->N + 127) / 128);
  tc_layer_kernel<<<grid, kTlThreads, kTlSmem, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// the persistent 128 x 256-tile form of the same layer (same parameters, same results bit for bit: the k order of the
// accumulation is unchanged); pays off from a few waves of row tiles on
int robir_tl_layer_big(const TcLayerParams* p, int sm_count, void* stream) {
  if (p->n == 0) return 0;
  RB_REQUIRE(p->nkb >= 1 && p->nkb <= 8 && p->N >= 1, "tl_layer_big: 1..8 k-blocks (K <= 512)");
  RB_CHECK_CUDA(cudaFuncSetAttribute(tc_layer_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTbSmem));
  const int col_tiles = ((p->N + 127) / 128 + 1) / 2, tiles = ((p->n + 127) / 128) * col_tiles;
  tc_layer_big_kernel<<<tiles < sm_count ? tiles : sm_count, kTlThreads, kTbSmem, (cudaStream_t)stream>>>(*p);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Weight gradients of one layer on the tensor cores (large row counts: the CESR shadow / normal networks).
//   G [n][ldg] (columns < N), A [n][lda] (columns < K), fp32  ->  dW [N][K] = G^T A, db [N] = column sums of G
// work: robir_tl_wgrad_workspace(n, N, K, sm_count) bytes (transposed hi/lo images of G and A, split partials, db
// partials).  Sums run in a fixed order: results are bitwise reproducible.  n_active (device, or null): rows at and
// beyond *n_active are zero (fixed-capacity batches) -- they are skipped, not read.
static void tl_wgrad_shape(int n, int N, int K, int sm_count, int* nkb, int* rt, int* cb, int* kps, int* splits) {
  *nkb = (n + 63) / 64;
  *rt = (N + 127) / 128;
  *cb = (K + 127) / 128;
  int want = sm_count / (*rt * *cb);
  if (want < 1) want = 1;
  if (want > *nkb) want = *nkb;
  *kps = (*nkb + want - 1) / want;
  *splits = (*nkb + *kps - 1) / *kps;
}

long long robir_tl_wgrad_workspace(int n, int N, int K, int sm_count) {
  int nkb, rt, cb, kps, splits;
  tl_wgrad_shape(n, N, K, sm_count, &nkb, &rt, &cb, &kps, &splits);
  return (long long)(rt + cb) * nkb * kTlBlockBytes + (long long)splits * rt * 128 * cb * 128 * 4 +
         (long long)nkb * rt * 128 * 4;
}

int robir_tl_wgrad(const float* G, int ldg, const float* A, int lda, int n, int N, int K, const int* n_active,
                   void* work, float* dW, float* db, int sm_count, void* stream) {
  RB_REQUIRE(n >= 1 && N >= 1 && K >= 1 && work != nullptr && dW != nullptr, "tl_wgrad: empty problem or no workspace");
  int nkb, rt, cb, kps, splits;
  tl_wgrad_shape(n, N, K, sm_count, &nkb, &rt, &cb, &kps, &splits);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* g_img = (uint8_t*)work;
  uint8_t* a_img = g_img + (size_t)rt * nkb * kTlBlockBytes;
  float* partial = (float*)(a_img + (size_t)cb * nkb * kTlBlockBytes);
  float* colsum = partial + (size_t)splits * rt * 128 * cb * 128;
  tl_pack_rows_t_kernel<<<dim3(nkb, rt), 256, 0, st>>>(G, ldg, n, N, nkb, g_img, db ? colsum : nullptr, rt * 128,
                                                       n_active);
  tl_pack_rows_t_kernel<<<dim3(nkb, cb), 256, 0, st>>>(A, lda, n, K, nkb, a_img, nullptr, 0, n_active);
  TlWgradParams p;
  p.g_img = g_img; p.a_img = a_img; p.nkb = nkb; p.kb_per_split = kps; p.row_tiles = rt; p.col_blocks = cb;
  p.partial = partial;
  p.n_active = n_active;
  RB_CHECK_CUDA(cudaFuncSetAttribute(tl_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTlSmem));
  tl_wgrad_kernel<<<dim3(rt * cb, splits), 192, kTlSmem, st>>>(p);
  const int dw_blocks = (int)(((long long)N * K + 255) / 256), db_blocks = db ? (N + 7) / 8 : 0;
  tl_wgrad_reduce_kernel<<<dw_blocks + db_blocks, 256, 0, st>>>(partial, splits, rt * 128, cb * 128, N, K, dW, colsum, nkb,
                                                                rt * 128, db, dw_blocks, n_active);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// The same weight gradients from the layer engine's images (see tl_wgrad_mn_kernel): g_img = image of G [n][N] (nkb_g
// k-blocks per row tile), a_img = image of A [n][K]; rows of G at and beyond *n_active (or n) must be zero in the image.
// work: robir_tl_wgrad_mn_workspace(N, K, sm_count) bytes.
static void tl_wgrad_mn_shape(int n, int N, int K, int sm_count, int* tn, int* tk, int* splits) {
  *tn = (N + 127) / 128;
  *tk = (K + 127) / 128;
  const int tiles = (n + 127) / 128;
  int want = sm_count / (*tn * *tk);
  if (want < 1) want = 1;
  *splits = want < tiles ? want : tiles;
}

long long robir_tl_wgrad_mn_workspace(int n, int N, int K, int sm_count) {
  int tn, tk, splits;
  tl_wgrad_mn_shape(n, N, K, sm_count, &tn, &tk, &splits);
  return (long long)splits * tn * 128 * tk * 128 * 4 + (long long)splits * tn * 128 * 4;
}

int robir_tl_wgrad_mn(const void* g_img, int nkb_g, const void* a_img, int nkb_a, int n, int N, int K,
                      const int* n_active, void* work, float* dW, float* db, int sm_count, void* stream) {
  RB_REQUIRE(n >= 1 && N >= 1 && K >= 1 && work != nullptr && dW != nullptr, "tl_wgrad_mn: empty problem or no workspace");
  RB_REQUIRE(nkb_g * 64 >= N && nkb_a * 64 >= K, "tl_wgrad_mn: images narrower than the matrices");
  int tn, tk, splits;
  tl_wgrad_mn_shape(n, N, K, sm_count, &tn, &tk, &splits);
  cudaStream_t st = (cudaStream_t)stream;
  TlWgradMnParams p;
  p.g_img = (const uint8_t*)g_img; p.a_img = (const uint8_t*)a_img; p.nkb_g = nkb_g; p.nkb_a = nkb_a;
  p.row_tiles = (n + 127) / 128; p.n = n; p.tiles_n = tn; p.tiles_k = tk;
  p.partial = (float*)work;
  p.db_part = db ? p.partial + (size_t)splits * tn * 128 * tk * 128 : nullptr;
  p.n_active = n_active;
  const int smem = kTlStages * kTlStageBytes + 1024 + 2048;
  RB_CHECK_CUDA(cudaFuncSetAttribute(tl_wgrad_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tl_wgrad_mn_kernel<<<dim3(tn * tk, splits), 192, smem, st>>>(p);
  const int dw_blocks = (int)(((long long)N * K + 255) / 256), db_blocks = db ? (N + 7) / 8 : 0;
  tl_wgrad_reduce_kernel<<<dw_blocks + db_blocks, 256, 0, st>>>(p.partial, splits, tn * 128, tk * 128, N, K, dW, p.db_part,
                                                                splits, tn * 128, db, dw_blocks, nullptr);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
