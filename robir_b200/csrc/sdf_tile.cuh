// Shared pieces of the NeuS SDF network evaluation (model/neus_model.py:312-438): PE rows and a CTA-level
// "evaluate the SDF of R points held in shared memory" routine, used inline by the ray-march kernels
// (sphere_trace.cu) so that a march step never leaves the kernel.
#pragma once
#include "mlp_engine.cuh"

namespace robir {

// write PE10 (value row) and its three directional derivatives (tangent rows) into tile rows [k0, k0+63)
template <bool JET>
__device__ __forceinline__ void sdf_pe_rows(float* Xs, int RP, int k0, int pt_local, const float* x, bool valid,
                                            float scale) {
  const int rbase = JET ? pt_local * 4 : pt_local;
  for (int i = 0; i < 3; ++i) {
    Xs[(k0 + i) * RP + rbase] = valid ? x[i] * scale : 0.f;
    if (JET)
      for (int j = 0; j < 3; ++j) Xs[(k0 + i) * RP + rbase + 1 + j] = (valid && i == j) ? scale : 0.f;
  }
  float f = 1.f;
  for (int l = 0; l < 10; ++l) {
    for (int i = 0; i < 3; ++i) {
      float sn = 0.f, cs = 0.f;
      if (valid) {
        sn = sinf(x[i] * f);
        cs = cosf(x[i] * f);
      }
      Xs[(k0 + 3 + 6 * l + i) * RP + rbase] = sn * scale;
      Xs[(k0 + 6 + 6 * l + i) * RP + rbase] = cs * scale;
      if (JET)
        for (int j = 0; j < 3; ++j) {
          Xs[(k0 + 3 + 6 * l + i) * RP + rbase + 1 + j] = (i == j) ? f * cs * scale : 0.f;
          Xs[(k0 + 6 + 6 * l + i) * RP + rbase + 1 + j] = (i == j) ? -f * sn * scale : 0.f;
        }
    }
    f *= 2.f;
  }
}


// Folded SDF network weights (ops.SdfWeights): layers 0..7 packed [Kpad][256] (layer 0: K = 64; layer 3: 193 valid
// columns), layer 8 row 0 kept as a vector.
struct SdfNet {
  const float* Wt[8];
  const float* bias[8];
  const float* w8_sdf;  // [256]
  const float* b8;      // [257] (only b8[0] is read here)
};

// s_out[r] = net(s_x[r])[0] for the R rows of the tile (0 for rows with s_valid[r] == 0).  s_x holds NETWORK-space
// points (already multiplied by the input scale).  All 256 threads must call; Xs: [256][R+4] floats, Wbuf:
// kWbufFloats, red: 256 floats.  Ends with a __syncthreads().
template <int R>
__device__ __forceinline__ void sdf_tile_values(const SdfNet& W, float* Xs, float* Wbuf, float* red,
                                                const float (*s_x)[3], const unsigned char* s_valid, float* s_out) {
  constexpr int RP = TileCfg<R>::RP, TR = TileCfg<R>::TR;
  const int tid = threadIdx.x, lane = tid & 31;
  const float kInvSqrt2 = 0.70710678118654752440f;
  if (tid < R) {
    sdf_pe_rows<false>(Xs, RP, 0, tid, s_x[tid], s_valid[tid] != 0, 1.f);
    Xs[63 * RP + tid] = 0.f;
  }
  __syncthreads();
  float acc[TR][8];
  for (int layer = 0; layer < 8; ++layer) {
    zero_acc<R>(acc);
    tile_gemm_pass<R>(Xs, layer == 0 ? 64 : 256, W.Wt[layer], 256, 0, Wbuf, acc);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(W.bias[layer] + lane * 4));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(W.bias[layer] + 128 + lane * 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float post = layer == 3 ? kInvSqrt2 : 1.f;   // x = cat([x, pe]) / sqrt(2) before layer 4
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int r = 0; r < TR; ++r) acc[r][c] = softplus100(acc[r][c] + bb[c]) * post;
    store_acc<R>(Xs, 0, layer == 3 ? 193 : 256, acc);
    if (layer == 3) {
      __syncthreads();
      if (tid < R) sdf_pe_rows<false>(Xs, RP, 193, tid, s_x[tid], s_valid[tid] != 0, kInvSqrt2);
    }
    __syncthreads();
  }
  {
    constexpr int PARTS = 256 / R, KLEN = 256 / PARTS;
    const int row = tid % R, part = tid / R;
    float s = 0.f;
    for (int k = part * KLEN; k < part * KLEN + KLEN; ++k) s = fmaf(__ldg(W.w8_sdf + k), Xs[k * RP + row], s);
    red[part * R + row] = s;
  }
  __syncthreads();
  if (tid < R) {
    constexpr int PARTS = 256 / R;
    float d = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) d += red[q * R + tid];
    s_out[tid] = s_valid[tid] ? d + __ldg(W.b8) : 0.f;
  }
  __syncthreads();
}

}  // namespace robir
