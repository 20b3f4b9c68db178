// IDR sphere tracer: RayTracing.forward / sphere_tracing / ray_sampler / secant /
// minimal_sdf_points of model/ray_tracing.py:26-326 and rend_util.get_sphere_intersection (utils/rend_util.py:141-163).
//
// The reference runs ~45 batch-wide SDF calls with boolean-mask indexing (a host sync each).  Here every march phase is
// a persistent kernel that evaluates the NeuS SDF network INLINE (sdf_tile_values: PE -> 9 layers on a tile of rows in
// shared memory), so a ray's whole march never leaves the SM:
//   1. sphere_trace_kernel   one tile = R/2 rays (start + end point rows); <= 10 two-sided steps, each with <= 3
//                            back-off halvings, loop control by __syncthreads_or; unconverged rays are appended to the
//                            sampler list (device counter, no host sync);
//   2. sample_eval_kernel    n_steps samples per listed ray, generated on the fly (mode 0: linspace in [t0, t1];
//                            mode 1: the caller's uniform offsets, training-only minimal_sdf_points);
//   3. sample_select_kernel  first sign change / arg-min sample, secant work list;
//   4. secant_kernel         n_secant_steps bracketing iterations, one tile = R rays;
//   5. minsdf_list / minsdf_select  (self.training only) closest-approach points for non-hit rays.
// Per-ray results do not depend on which other rays share the batch (every global loop of the reference is inert for
// finished rays), with one exception that is reproduced through a device flag: rays that miss the bounding sphere
// return the camera origin iff the global march loop ran at least once (ray_tracing.py:163-164).
#include "sdf_tile.cuh"

namespace robir {

struct SphereTraceParams {
  SdfNet net;
  int N, o_div;
  const float* cam_loc;               // [N / o_div][3]
  const float* ray_dirs;              // [N][3]
  const unsigned char* object_mask;   // [N] or null (= all true)
  float in_scale, out_scale;          // f(p) = net(in_scale * p)[0] * out_scale
  float radius, sdf_threshold, line_search_step;
  int line_step_iters, sphere_tracing_iters, n_steps, n_secant_steps, training;
  const float* uniform_steps;         // [n_steps], training only
  float* points;                      // [N][3]
  unsigned char* net_mask;            // [N]
  float* dists;                       // [N]
  // caller-provided workspace (robir_sphere_trace_workspace_bytes)
  float* acc_s;                       // [N]
  float* acc_e;                       // [N]
  float* min_dis;                     // [N]
  float* max_dis;                     // [N]
  unsigned char* flags;               // [N] bit0 = hits the bounding sphere, bit1 = sampler ray
  int* samp_list;                     // [N]
  int* sec_list;                      // [N]
  float* sec_state;                   // [N][4]  z_lo, f_lo, z_hi, f_hi
  int* min_list;                      // [N]
  float* vals;                        // [N * n_steps]
  int* counters;                      // [8] zero-initialised: 0 n_samp, 1 n_sec, 2 n_min, 3 loop_ran, 4 sdf queries
};

enum { CNT_SAMP = 0, CNT_SEC = 1, CNT_MIN = 2, CNT_LOOP = 3, CNT_QUERIES = 4 };

__device__ __forceinline__ void ray_origin(const SphereTraceParams& p, int ray, float o[3], float d[3]) {
  const int io = ray / p.o_div;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = p.cam_loc[3 * io + c];
    d[c] = p.ray_dirs[3 * (size_t)ray + c];
  }
}
// torch evaluates o + t * d as a rounded product followed by a rounded sum
__device__ __forceinline__ float madd(float o, float t, float d) { return __fadd_rn(o, __fmul_rn(t, d)); }

// torch.linspace(0, 1, n)[i] in fp32 (symmetric evaluation: start + i*step below the midpoint, end - (n-1-i)*step above)
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n == 1) return 0.f;
  const float step = 1.f / (float)(n - 1);
  return i < n / 2 ? __fmul_rn(step, (float)i) : __fsub_rn(1.f, __fmul_rn(step, (float)(n - 1 - i)));
}

template <int R>
struct TraceSmem {
  float x[R][3];
  unsigned char valid[R];
  float out[R];
};

// ---------------------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256, 2) sphere_trace_kernel(SphereTraceParams p) {
  constexpr int H = R / 2, RP = TileCfg<R>::RP;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Wbuf = Xs + 256 * RP;
  float* red = Wbuf + kWbufFloats;
  __shared__ TraceSmem<R> s;
  const int tid = threadIdx.x;
  const int ntile = (p.N + H - 1) / H;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int ray = tile * H + tid;
    const bool mine = tid < H && ray < p.N;
    float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
    bool hit_sphere = false, un_s = false, un_e = false;
    float acc_s = 0.f, acc_e = 0.f, cur_s = 0.f, cur_e = 0.f, nxt_s = 0.f, nxt_e = 0.f, min_d = 0.f, max_d = 0.f;
    int n_q = 0;
    if (mine) {
      ray_origin(p, ray, o, d);
      // get_sphere_intersection: b = d.o, under = b^2 - (|o|^2 - r^2)
      const float b = d[0] * o[0] + d[1] * o[1] + d[2] * o[2];
      const float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
      const float under = b * b - (nrm * nrm - p.radius * p.radius);
      hit_sphere = under > 0.f;
      if (hit_sphere) {
        const float sq = sqrtf(under);
        acc_s = fmaxf(-sq - b, 0.01f);
        acc_e = fmaxf(sq - b, 0.01f);
      }
      un_s = un_e = hit_sphere;
      min_d = acc_s;
      max_d = acc_e;
    }
    auto stage = [&](bool vs, bool ve) {   // stage the start / end points of this thread's ray as rows tid, H + tid
      if (tid < H) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s.x[tid][c] = madd(o[c], acc_s, d[c]) * p.in_scale;
          s.x[H + tid][c] = madd(o[c], acc_e, d[c]) * p.in_scale;
        }
        s.valid[tid] = vs;
        s.valid[H + tid] = ve;
        n_q += (int)vs + (int)ve;
      }
      __syncthreads();
      sdf_tile_values<R>(p.net, Xs, Wbuf, red, s.x, s.valid, s.out);
    };
    stage(un_s, un_e);
    if (tid < H) {
      nxt_s = un_s ? s.out[tid] * p.out_scale : 0.f;
      nxt_e = un_e ? s.out[H + tid] * p.out_scale : 0.f;
    }
    int it = 0;
    while (true) {
      cur_s = un_s ? nxt_s : 0.f;
      if (cur_s <= p.sdf_threshold) cur_s = 0.f;
      cur_e = un_e ? nxt_e : 0.f;
      if (cur_e <= p.sdf_threshold) cur_e = 0.f;
      un_s = un_s && cur_s > p.sdf_threshold;
      un_e = un_e && cur_e > p.sdf_threshold;
      const int any = __syncthreads_or(un_s || un_e);
      if (!any || it == p.sphere_tracing_iters) break;
      ++it;
      if (tid == 0 && it == 1) atomicExch(p.counters + CNT_LOOP, 1);
      acc_s = __fadd_rn(acc_s, cur_s);
      acc_e = __fsub_rn(acc_e, cur_e);
      stage(un_s, un_e);
      if (tid < H) {
        nxt_s = un_s ? s.out[tid] * p.out_scale : 0.f;
        nxt_e = un_e ? s.out[H + tid] * p.out_scale : 0.f;
      }
      bool bad_s = nxt_s < 0.f, bad_e = nxt_e < 0.f;
      int back = 0;
      while (back < p.line_step_iters && __syncthreads_or(bad_s || bad_e)) {
        const float f = (1.f - p.line_search_step) / (float)(1 << back);
        if (bad_s) acc_s = __fsub_rn(acc_s, __fmul_rn(f, cur_s));
        if (bad_e) acc_e = __fadd_rn(acc_e, __fmul_rn(f, cur_e));
        stage(bad_s, bad_e);
        if (tid < H) {
          if (bad_s) nxt_s = s.out[tid] * p.out_scale;
          if (bad_e) nxt_e = s.out[H + tid] * p.out_scale;
        }
        bad_s = nxt_s < 0.f;
        bad_e = nxt_e < 0.f;
        ++back;
      }
      un_s = un_s && acc_s < acc_e;
      un_e = un_e && acc_s < acc_e;
    }
    if (mine) {
      p.acc_s[ray] = acc_s;
      p.acc_e[ray] = acc_e;
      p.min_dis[ray] = min_d;
      p.max_dis[ray] = max_d;
      p.flags[ray] = (unsigned char)((hit_sphere ? 1 : 0) | (un_s ? 2 : 0));
      p.net_mask[ray] = acc_s < acc_e;
      p.dists[ray] = acc_s;
#pragma unroll
      for (int c = 0; c < 3; ++c) p.points[3 * (size_t)ray + c] = madd(o[c], acc_s, d[c]);
      if (un_s) p.samp_list[atomicAdd(p.counters + CNT_SAMP, 1)] = ray;
      if (n_q) atomicAdd(p.counters + CNT_QUERIES, n_q);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// vals[j * n_steps + i] = f(o + t_i * d) for the j-th listed ray.  mode 0: sampler (ray_tracing.py:208-225);
// mode 1: minimal_sdf_points (:299-317).
template <int R>
__global__ void __launch_bounds__(256, 2) sample_eval_kernel(SphereTraceParams p, int mode) {
  constexpr int RP = TileCfg<R>::RP;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Wbuf = Xs + 256 * RP;
  float* red = Wbuf + kWbufFloats;
  __shared__ TraceSmem<R> s;
  const int tid = threadIdx.x;
  const int count = p.counters[mode == 0 ? CNT_SAMP : CNT_MIN];
  const int* list = mode == 0 ? p.samp_list : p.min_list;
  const long long rows = (long long)count * p.n_steps;
  const int ntile = (int)((rows + R - 1) / R);
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const long long r = (long long)tile * R + tid;
    if (tid < R) {
      const bool ok = r < rows;
      s.valid[tid] = ok;
      s.x[tid][0] = s.x[tid][1] = s.x[tid][2] = 0.f;
      if (ok) {
        const int j = (int)(r / p.n_steps), i = (int)(r % p.n_steps);
        const int ray = list[j];
        float o[3], d[3];
        ray_origin(p, ray, o, d);
        float t;
        if (mode == 0) {
          const float a = p.acc_s[ray], e = p.acc_e[ray];
          t = __fadd_rn(a, __fmul_rn(linspace01(i, p.n_steps), __fsub_rn(e, a)));
        } else {
          const float lo = p.min_dis[ray], hi = p.max_dis[ray];
          t = __fadd_rn(__fmul_rn(p.uniform_steps[i], __fsub_rn(hi, lo)), lo);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) s.x[tid][c] = madd(o[c], t, d[c]) * p.in_scale;
      }
    }
    __syncthreads();
    sdf_tile_values<R>(p.net, Xs, Wbuf, red, s.x, s.valid, s.out);
    if (tid < R && r < rows) p.vals[r] = s.out[tid] * p.out_scale;
    if (tid == 0) atomicAdd(p.counters + CNT_QUERIES, (int)min((long long)R, rows - (long long)tile * R));
    __syncthreads();
  }
}

// one thread per sampler ray: first sign change, fallback arg-min, secant bracket (ray_tracing.py:227-274)
__global__ void sample_select_kernel(SphereTraceParams p) {
  const int count = p.counters[CNT_SAMP];
  const int n = p.n_steps;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
    const int ray = p.samp_list[j];
    const float* v = p.vals + (size_t)j * n;
    // argmin_i sign(v_i) * (n - i): the first negative sample; else the first zero; else the last sample
    int first = -1, first_zero = -1, amin = 0;
    float vmin = v[0];
    for (int i = 0; i < n; ++i) {
      const float x = v[i];
      if (x < 0.f && first < 0) first = i;
      if (x == 0.f && first_zero < 0) first_zero = i;
      if (x < vmin) { vmin = x; amin = i; }
    }
    if (first < 0) first = first_zero >= 0 ? first_zero : n - 1;
    const float a = p.acc_s[ray], e = p.acc_e[ray];
    auto z_of = [&](int i) { return __fadd_rn(a, __fmul_rn(linspace01(i, n), __fsub_rn(e, a))); };
    const bool true_surf = p.object_mask ? p.object_mask[ray] != 0 : true;
    const bool net_surf = v[first] < 0.f;
    float o[3], d[3];
    ray_origin(p, ray, o, d);
    const int pick = (true_surf && net_surf) ? first : amin;
    const float z = z_of(pick);
    p.dists[ray] = z;
#pragma unroll
    for (int c = 0; c < 3; ++c) p.points[3 * (size_t)ray + c] = madd(o[c], z, d[c]);
    p.net_mask[ray] = net_surf;
    const bool sec = p.training ? (net_surf && true_surf) : net_surf;
    if (sec) {
      const int lo = (first - 1 + n) % n;   // python's negative index wraps to the last sample when first == 0
      const int k = atomicAdd(p.counters + CNT_SEC, 1);
      p.sec_list[k] = ray;
      p.sec_state[4 * k + 0] = z_of(lo);
      p.sec_state[4 * k + 1] = v[lo];
      p.sec_state[4 * k + 2] = z_of(first);
      p.sec_state[4 * k + 3] = v[first];
    }
  }
}

// secant root finding (ray_tracing.py:276-297): one tile = R rays, n_secant_steps inline SDF evaluations
template <int R>
__global__ void __launch_bounds__(256, 2) secant_kernel(SphereTraceParams p) {
  constexpr int RP = TileCfg<R>::RP;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Wbuf = Xs + 256 * RP;
  float* red = Wbuf + kWbufFloats;
  __shared__ TraceSmem<R> s;
  const int tid = threadIdx.x;
  const int count = p.counters[CNT_SEC];
  const int ntile = (count + R - 1) / R;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int k = tile * R + tid;
    const bool mine = tid < R && k < count;
    float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f}, z_lo = 0.f, f_lo = 0.f, z_hi = 0.f, f_hi = 0.f, zp = 0.f;
    int ray = 0;
    auto predict = [&]() {
      const float num = __fmul_rn(-f_lo, __fsub_rn(z_hi, z_lo));
      const float den = __fadd_rn(__fsub_rn(f_hi, f_lo), 1e-8f);
      return fminf(fmaxf(__fadd_rn(__fdiv_rn(num, den), z_lo), 0.f), 20.f);
    };
    if (mine) {
      ray = p.sec_list[k];
      ray_origin(p, ray, o, d);
      z_lo = p.sec_state[4 * k + 0];
      f_lo = p.sec_state[4 * k + 1];
      z_hi = p.sec_state[4 * k + 2];
      f_hi = p.sec_state[4 * k + 3];
      zp = predict();
    }
    for (int itr = 0; itr < p.n_secant_steps; ++itr) {
      if (tid < R) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s.x[tid][c] = madd(o[c], zp, d[c]) * p.in_scale;
        s.valid[tid] = mine;
      }
      __syncthreads();
      sdf_tile_values<R>(p.net, Xs, Wbuf, red, s.x, s.valid, s.out);
      if (mine) {
        const float fm = s.out[tid] * p.out_scale;
        if (fm > 0.f) { z_lo = zp; f_lo = fm; }
        if (fm < 0.f) { z_hi = zp; f_hi = fm; }
        zp = predict();
      }
    }
    if (mine) {
      p.dists[ray] = zp;
#pragma unroll
      for (int c = 0; c < 3; ++c) p.points[3 * (size_t)ray + c] = madd(o[c], zp, d[c]);
    }
    if (tid == 0) atomicAdd(p.counters + CNT_QUERIES, min(R, count - tile * R) * p.n_secant_steps);
    __syncthreads();
  }
}

// eval mode: rays that miss the bounding sphere keep p = origin (if the march loop ran) / 0, t = 0.
// training mode (ray_tracing.py:73-100): closest-approach point for rays that miss the sphere, work list for
// minimal_sdf_points for the other non-hit rays.
__global__ void finalize_kernel(SphereTraceParams p) {
  const int loop_ran = p.counters[CNT_LOOP];
  for (int ray = blockIdx.x * blockDim.x + threadIdx.x; ray < p.N; ray += gridDim.x * blockDim.x) {
    const unsigned char fl = p.flags[ray];
    const bool hit_sphere = fl & 1, samp = fl & 2;
    float o[3], d[3];
    ray_origin(p, ray, o, d);
    if (!hit_sphere && !loop_ran) {
#pragma unroll
      for (int c = 0; c < 3; ++c) p.points[3 * (size_t)ray + c] = 0.f;
    }
    if (!p.training) continue;
    const bool om = p.object_mask ? p.object_mask[ray] != 0 : true;
    const bool net = p.net_mask[ray] != 0;
    const bool in_mask = !net && om && !samp;
    const bool out_mask = !om && !samp;
    if (!(in_mask || out_mask)) continue;
    if (!hit_sphere) {
      const float t = -(__fadd_rn(__fadd_rn(__fmul_rn(d[0], o[0]), __fmul_rn(d[1], o[1])), __fmul_rn(d[2], o[2])));
      p.dists[ray] = t;
#pragma unroll
      for (int c = 0; c < 3; ++c) p.points[3 * (size_t)ray + c] = madd(o[c], t, d[c]);
    } else {
      if (net && out_mask) p.min_dis[ray] = p.dists[ray];
      p.min_list[atomicAdd(p.counters + CNT_MIN, 1)] = ray;
    }
  }
}

__global__ void minsdf_select_kernel(SphereTraceParams p) {
  const int count = p.counters[CNT_MIN];
  const int n = p.n_steps;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
    const int ray = p.min_list[j];
    const float* v = p.vals + (size_t)j * n;
    int amin = 0;
    float vmin = v[0];
    for (int i = 1; i < n; ++i)
      if (v[i] < vmin) { vmin = v[i]; amin = i; }
    const float lo = p.min_dis[ray], hi = p.max_dis[ray];
    const float t = __fadd_rn(__fmul_rn(p.uniform_steps[amin], __fsub_rn(hi, lo)), lo);
    float o[3], d[3];
    ray_origin(p, ray, o, d);
    p.dists[ray] = t;
#pragma unroll
    for (int c = 0; c < 3; ++c) p.points[3 * (size_t)ray + c] = madd(o[c], t, d[c]);
  }
}

template <int R>
static int launch_sphere_trace(const SphereTraceParams& p, int sm_count, cudaStream_t st) {
  const int smem = (256 * (R + 4) + kWbufFloats + 256) * 4;
  RB_CHECK_CUDA(cudaFuncSetAttribute(sphere_trace_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RB_CHECK_CUDA(cudaFuncSetAttribute(sample_eval_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  RB_CHECK_CUDA(cudaFuncSetAttribute(secant_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int cap = 2 * sm_count;
  const int tiles = (p.N + R / 2 - 1) / (R / 2);
  sphere_trace_kernel<R><<<tiles < cap ? tiles : cap, 256, smem, st>>>(p);
  const long long rows = (long long)p.N * p.n_steps;
  const int stiles = (int)((rows + R - 1) / R);
  sample_eval_kernel<R><<<stiles < cap ? stiles : cap, 256, smem, st>>>(p, 0);
  sample_select_kernel<<<(p.N + 127) / 128, 128, 0, st>>>(p);
  const int ctiles = (p.N + R - 1) / R;
  secant_kernel<R><<<ctiles < cap ? ctiles : cap, 256, smem, st>>>(p);   // 0 steps still applies the first prediction
  finalize_kernel<<<(p.N + 127) / 128, 128, 0, st>>>(p);
  if (p.training) {
    sample_eval_kernel<R><<<stiles < cap ? stiles : cap, 256, smem, st>>>(p, 1);
    minsdf_select_kernel<<<(p.N + 127) / 128, 128, 0, st>>>(p);
  }
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace robir

using namespace robir;

extern "C" {

// number of kernels robir_sphere_trace launches (for launch accounting)
int robir_sphere_trace_launches(int training) { return 5 + (training ? 2 : 0); }

int robir_sphere_trace(const SphereTraceParams* p, int sm_count, void* stream) {
  if (p->N == 0) return 0;
  RB_REQUIRE(p->o_div >= 1 && p->N % p->o_div == 0, "sphere_trace: N must be a multiple of the rays-per-origin count");
  RB_REQUIRE(p->n_steps >= 1 && p->line_step_iters >= 0 && p->line_step_iters < 31, "sphere_trace: bad step counts");
  RB_REQUIRE(!p->training || p->uniform_steps != nullptr, "sphere_trace: training mode needs uniform_steps[n_steps]");
  // few rays: 32-row tiles (16 rays) put more SMs to work on the latency-bound march; many rays: 64-row tiles
  if ((long long)p->N * 2 / 64 >= 2LL * sm_count) return launch_sphere_trace<64>(*p, sm_count, (cudaStream_t)stream);
  return launch_sphere_trace<32>(*p, sm_count, (cudaStream_t)stream);
}

}  // extern "C"
