// CTA-level fp32 tile-GEMM engine (FFMA path).  A tile of R activation rows lives in shared memory, K-major
// (X[k][row], row stride RP = R + 4 floats), weights stream from global/L2 through a double-buffered cp.async ring in
// 16 x 256 chunks.  Used by: the fused visibility MLP (forward + input-gradient), the SDF value/normal kernel and the
// small material/indirect networks.  (The tcgen05 tensor-core engine for the visibility MLP is in vis_tc.cu.)
//
// Thread mapping (256 threads): warp w owns rows [w*R/8, (w+1)*R/8); lane l owns columns {4l..4l+3} and {128+4l..+3}
// of the current 256-column pass.  The same mapping is used for the ReLU bit masks: word (g*4+j) of a row holds, at
// bit l, the sign of column g*128 + 4l + j.
#pragma once
#include "common.cuh"

namespace robir {

constexpr int kPassCols = 256;
constexpr int kChunkK = 16;
constexpr int kWbufFloats = 2 * kChunkK * kPassCols;  // 32 KB

template <int R>
struct TileCfg {
  static constexpr int RP = R + 4;
  static constexpr int TR = R / 8;
};

__device__ __forceinline__ int tile_col(int lane, int c) { return c < 4 ? lane * 4 + c : 128 + lane * 4 + (c - 4); }

// acc[r][c] (+)= sum_k Xs[k][row0+r] * Wt[k][col0 + tile_col(c)];   K % 16 == 0; Wt rows are ldw floats apart
// (ldw % 4 == 0, col0 % 4 == 0, buffer padded so that [K][col0 .. col0+255] is readable).
// Ends with a __syncthreads(): on return no thread still reads Xs or Wbuf.
template <int R>
__device__ __forceinline__ void tile_gemm_pass(const float* __restrict__ Xs, int K, const float* __restrict__ Wt,
                                               int ldw, int col0, float* __restrict__ Wbuf,
                                               float (&acc)[R / 8][8]) {
  constexpr int RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const float* wb = Wbuf + (c & 1) * (kChunkK * kPassCols);
    const float* xb = Xs + (size_t)(c * kChunkK) * RP + warp * TR;
#pragma unroll
    for (int kk = 0; kk < kChunkK; ++kk) {
      float xr[TR];
#pragma unroll
      for (int i = 0; i < TR; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xb + kk * RP + i);
        xr[i] = v.x; xr[i + 1] = v.y; xr[i + 2] = v.z; xr[i + 3] = v.w;
      }
      const float4 w0 = *reinterpret_cast<const float4*>(wb + kk * kPassCols + lane * 4);
      const float4 w1 = *reinterpret_cast<const float4*>(wb + kk * kPassCols + 128 + lane * 4);
      const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) acc[r][cc] = fmaf(xr[r], wr[cc], acc[r][cc]);
    }
    __syncthreads();
  }
}

template <int R>
__device__ __forceinline__ void zero_acc(float (&acc)[R / 8][8]) {
#pragma unroll
  for (int r = 0; r < R / 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
}

// Store a thread's accumulator block back into the K-major tile: Xs[col][row0 + r] = acc[r][c]
template <int R>
__device__ __forceinline__ void store_acc(float* __restrict__ Xs, int col0, int ncols, const float (&acc)[R / 8][8]) {
  constexpr int RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int col = col0 + tile_col(lane, c);
    if (col < ncols) {
      float* dst = Xs + (size_t)col * RP + warp * TR;
#pragma unroll
      for (int i = 0; i < TR; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i][c], acc[i + 1][c], acc[i + 2][c], acc[i + 3][c]);
    }
  }
}

// ReLU sign bits of a thread block's accumulators in the ballot layout described above.
// mask_row(r) must return the address of the 8-word record of local row r (or nullptr to skip).
template <int R, typename F>
__device__ __forceinline__ void store_relu_mask(const float (&acc)[R / 8][8], F mask_row) {
  constexpr int TR = TileCfg<R>::TR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < TR; ++r) {
    unsigned mine = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const unsigned b = __ballot_sync(0xffffffffu, acc[r][c] > 0.f);
      if (lane == c) mine = b;
    }
    uint32_t* dst = mask_row(warp * TR + r);
    if (dst != nullptr && lane < 8) dst[lane] = mine;
  }
}

template <int R>
__device__ __forceinline__ void apply_relu_mask(float (&acc)[R / 8][8], const uint32_t* const (&rows)[R / 8]) {
  constexpr int TR = TileCfg<R>::TR;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < TR; ++r) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t w = rows[r] ? __ldg(rows[r] + c) : 0u;
      if (!((w >> lane) & 1u)) acc[r][c] = 0.f;
    }
  }
}

}  // namespace robir
