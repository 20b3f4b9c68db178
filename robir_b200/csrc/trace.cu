// Camera rays and the octree surface tracer (row a4): one cooperative, lock-step kernel replaces
// the ~5000 ATen launches + host syncs of OctreeSDF.cast.  Reference: utils/rend_util.py:51-97 (get_camera_params,
// lift), utils/octree.py:421-438,459-471,493-585, model/octree_tracing.py:43-60.
// Compiled with -fmad=false: every product/sum is rounded separately like the reference's elementwise ops.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "octree_walk.h"

namespace cg = cooperative_groups;

namespace robir {

// uv [N][2], pose [4][4] row-major (only rows 0..2 used), K [3][3] -> dirs [N][3]; cam_loc = pose[:3,3]
__global__ void camera_rays_kernel(int N, const float* __restrict__ uv, const float* __restrict__ pose,
                                   const float* __restrict__ Kmat, float* __restrict__ dirs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float fx = Kmat[0], sk = Kmat[1], cx = Kmat[2], fy = Kmat[4], cy = Kmat[5];
  const float x = uv[2 * i], y = uv[2 * i + 1];
  const float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * 1.0f;
  const float yl = (y - cy) / fy * 1.0f;
  const float pc[4] = {xl, -yl, -1.0f, 1.0f};
  float w[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float s = pose[4 * r] * pc[0];
    s += pose[4 * r + 1] * pc[1];
    s += pose[4 * r + 2] * pc[2];
    s += pose[4 * r + 3] * pc[3];
    w[r] = s - pose[4 * r + 3];
  }
  const float nrm = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const float den = fmaxf(nrm, 1e-12f);  // F.normalize eps
  dirs[3 * i] = w[0] / den; dirs[3 * i + 1] = w[1] / den; dirs[3 * i + 2] = w[2] / den;
}

struct OctCastParams {
  OctreeView view;
  const float* sdf_grad;    // [n_nodes][3] unit gradients at node centres
  const float* rays_o;      // [K][3]  (or [Ko][3] with o_stride rays sharing one origin)
  const float* rays_d;      // [K][3]
  int K;
  int o_div;                // ray r uses origin r / o_div (1 = per-ray origins)
  int max_iter;             // -1: until all rays finish; >0: secondary-ray mode (bias 0.005, iteration cap)
  float eps;                // 1e-3
  float refine_limit;       // float32(10 * min_step)
  float last_node_sdf;      // sdf_val[-1]
  float* state_t;           // [K] workspace
  int* state_ptr;           // [K] workspace
  float* out_t;             // [K]
  float* out_x;             // [K][3]
  unsigned char* out_hit;   // [K]
  unsigned* counters;       // [kMaxIter + 8]: live count per iteration, then stats (zero-init by caller)
};
constexpr int kMaxIter = 4096;
// counters[kMaxIter + 0] = node visits (low), +1 micro samples, +2 iterations executed

// CLUSTER = false: cooperative grid of any size, lock-step through grid.sync() and global live counters.
// CLUSTER = true : ONE thread-block cluster (<= 16 CTAs x 1024 threads) holds every ray of the call: the per-iteration
// barrier is the hardware cluster barrier and the live count is summed through distributed shared memory -- the
// reference's multi_samp depends on that count every iteration (utils/octree.py:547-548), so the iterations cannot be
// decoupled, only the barrier made cheap.  Same per-ray arithmetic in the same order: bit-identical results.
template <bool CLUSTER>
__global__ void __launch_bounds__(CLUSTER ? 1024 : 256) octree_cast_kernel(OctCastParams p) {
  cg::grid_group grid = cg::this_grid();
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned s_live[3];
  if (CLUSTER && threadIdx.x < 3) s_live[threadIdx.x] = 0;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  const OctreeView& o = p.view;
  const bool secondary = p.max_iter > 0;
  unsigned visits = 0, samples = 0;

  auto origin = [&](int r, float* ox, float* oy, float* oz, float dx, float dy, float dz) {
    const float* po = p.rays_o + 3 * (size_t)(r / p.o_div);
    *ox = po[0]; *oy = po[1]; *oz = po[2];
    if (secondary) { *ox = *ox + dx * 0.005f; *oy = *oy + dy * 0.005f; *oz = *oz + dz * 0.005f; }
  };

  // ---- init: lanes over rays
  {
    unsigned live_cnt = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.K; r += gridDim.x * blockDim.x) {
      const float dx = p.rays_d[3 * r], dy = p.rays_d[3 * r + 1], dz = p.rays_d[3 * r + 2];
      float ox, oy, oz;
      origin(r, &ox, &oy, &oz, dx, dy, dz);
      RayState s;
      ray_init(o, ox, oy, oz, dx, dy, dz, p.eps, &s, &visits);
      p.state_t[r] = s.t;
      p.state_ptr[r] = s.live ? s.ptr : (s.ptr >= 0 ? -2 - s.ptr : -1);
      live_cnt += s.live ? 1u : 0u;
    }
    live_cnt = __reduce_add_sync(0xffffffffu, live_cnt);
    if (CLUSTER) {
      __syncthreads();                                   // s_live zeroed
      if (lane == 0 && live_cnt) atomicAdd(&s_live[0], live_cnt);
    } else if (lane == 0 && live_cnt) {
      atomicAdd(&p.counters[0], live_cnt);
    }
  }
  if (CLUSTER) cluster.sync(); else grid.sync();

  // state_ptr encoding: >= 0 live in node ptr;  -1 dead outside;  <= -2 finished (hit) in node (-2 - v)
  int it = 0;
  for (;; ++it) {
    unsigned live;
    if (CLUSTER) {
      // sum of every CTA's count of rays alive after iteration it - 1 (slot it % 3 of each CTA's shared memory)
      unsigned part = 0;
      if (lane < (int)cluster.num_blocks()) part = *cluster.map_shared_rank(&s_live[it % 3], lane);
      live = __reduce_add_sync(0xffffffffu, part);
      if (threadIdx.x == 0) s_live[(it + 2) % 3] = 0;     // read by everyone one iteration ago, next filled in it + 1
      if (blockIdx.x == 0 && threadIdx.x == 0) p.counters[it] = live;   // statistics only
    } else {
      live = *((volatile unsigned*)&p.counters[it]);
    }
    if (live == 0 || it >= kMaxIter - 1) break;
    if (secondary && it > p.max_iter) break;
    float step = 0.001f;
    if (secondary) step = p.K > 100000 ? 0.01f : 0.005f;
    long long q = (long long)p.K * 10;
    q = q < 1 ? 1 : (q > 2000000 ? 2000000 : q);
    int ms = (int)(q / (long long)live);
    ms = ms < 1 ? 1 : (ms > 100 ? 100 : ms);
    unsigned live_next = 0;
    for (int r = gwarp; r < p.K; r += nwarp) {       // one warp per ray: the micro-march runs across lanes
      const int ptr = p.state_ptr[r];
      if (ptr < 0) continue;
      const float dx = p.rays_d[3 * r], dy = p.rays_d[3 * r + 1], dz = p.rays_d[3 * r + 2];
      float ox, oy, oz;
      origin(r, &ox, &oy, &oz, dx, dy, dz);
      float t = p.state_t[r];
      const float px = ox + t * dx, py = oy + t * dy, pz = oz + t * dz;
      OctNode nd;
      RB_LDG_NODE(nd, o.nodes + ptr);
      float far = box_far(nd.minx, nd.miny, nd.minz, nd.sizex, nd.sizey, nd.sizez, px, py, pz, dx, dy, dz, nullptr,
                          nullptr);
      if (far < (float)ms * step) {
        // fast_volume_render across lanes: sample c at t[c+1]; first c with cached sdf <= step
        int first = ms;
        for (int c0 = 0; c0 < ms && first == ms; c0 += 32) {
          const int c = c0 + lane;
          bool hit = false;
          if (c < ms) {
            const float ts = linspace01(c + 1, ms + 1) * (float)ms * step + step;
            OctNode sn;
            unsigned v = 0;
            const int sp = oct_query(o, px + dx * ts, py + dy * ts, pz + dz * ts, &sn, &v);
            if (lane == 0 || true) visits += v;
            ++samples;
            const float sdf = sp >= 0 ? sn.sdf_val : p.last_node_sdf;
            hit = sdf <= step;
          }
          const unsigned b = __ballot_sync(0xffffffffu, hit);
          if (b) first = c0 + __ffs(b) - 1;
        }
        far = linspace01(first, ms + 1) * (float)ms * step + step;
      }
      t = t + (far + p.eps);
      const float nx = ox + t * dx, ny = oy + t * dy, nz = oz + t * dz;
      int nptr = -1;
      bool alive = false;
      if (inside_root_open(o, nx, ny, nz)) {
        OctNode nn;
        unsigned v = 0;
        nptr = oct_query(o, nx, ny, nz, &nn, &v);
        if (lane == 0) visits += v;
        alive = (nptr >= 0) && !node_is_hit(nn);
      }
      if (lane == 0) {
        p.state_t[r] = t;
        p.state_ptr[r] = alive ? nptr : (nptr >= 0 ? -2 - nptr : -1);
        live_next += alive ? 1u : 0u;
      }
    }
    if (CLUSTER) {
      if (lane == 0 && live_next) atomicAdd(&s_live[(it + 1) % 3], live_next);
      cluster.sync();
    } else {
      if (lane == 0 && live_next) atomicAdd(&p.counters[it + 1], live_next);
      grid.sync();
    }
  }

  // a CTA must not exit (and release its shared memory) while a slower CTA of the cluster may still be reading its
  // live count for the iteration that ended the loop
  if (CLUSTER) cluster.sync();

  // ---- finish: hit test + first-order refinement
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.K; r += gridDim.x * blockDim.x) {
    const int enc = p.state_ptr[r];
    const int ptr = enc >= 0 ? enc : (enc <= -2 ? -2 - enc : -1);
    const float dx = p.rays_d[3 * r], dy = p.rays_d[3 * r + 1], dz = p.rays_d[3 * r + 2];
    float ox, oy, oz;
    origin(r, &ox, &oy, &oz, dx, dy, dz);
    float t = p.state_t[r];
    if (ptr >= 0) {
      RayState s;
      s.t = t;
      s.px = ox + t * dx; s.py = oy + t * dy; s.pz = oz + t * dz;
      s.ptr = ptr;
      RB_LDG_NODE(s.node, o.nodes + ptr);
      t = refine_t(s, p.sdf_grad[3 * (size_t)ptr], p.sdf_grad[3 * (size_t)ptr + 1], p.sdf_grad[3 * (size_t)ptr + 2], dx,
                   dy, dz, p.refine_limit);
    }
    const float* po = p.rays_o + 3 * (size_t)(r / p.o_div);   // hit_x uses the UNBIASED origin (octree_tracing.py:56)
    p.out_t[r] = t;
    p.out_hit[r] = ptr >= 0 ? 1 : 0;
    p.out_x[3 * r] = t * dx + po[0]; p.out_x[3 * r + 1] = t * dy + po[1]; p.out_x[3 * r + 2] = t * dz + po[2];
  }
  visits = __reduce_add_sync(0xffffffffu, visits);
  samples = __reduce_add_sync(0xffffffffu, samples);
  if (lane == 0) {
    atomicAdd(&p.counters[kMaxIter], visits);
    atomicAdd(&p.counters[kMaxIter + 1], samples);
    if (gwarp == 0) p.counters[kMaxIter + 2] = (unsigned)it;
  }
}

}  // namespace robir

using namespace robir;

extern "C" {

int robir_camera_rays(int N, const float* uv, const float* pose, const float* K, float* dirs, void* stream) {
  if (N == 0) return 0;
  camera_rays_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, uv, pose, K, dirs);
  RB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int robir_octree_counters_len() { return kMaxIter + 8; }

// counters: uint32 [robir_octree_counters_len()], zero-initialised by the caller before every call.
int robir_octree_cast(const OctCastParams* p, int sm_count, void* stream) {
  if (p->K == 0) return 0;
  RB_REQUIRE(p->o_div >= 1, "octree_cast: o_div must be >= 1");
  OctCastParams params = *p;
  // ---- experiment (ROBIR_OCTREE_CLUSTER=1): one thread-block cluster, hardware barrier per iteration.  Measured on the
  // bench batch (1024 rays, ~55 iterations): 560 us vs 365 us for the cooperative grid -- a cluster is at most 16 SMs
  // (512 warps: two rays per warp, walked one after the other), and an iteration is bound by the ~15 dependent L2 round
  // trips of one ray step (~4.5 us), not by the grid barrier (~2 us).  Kept selectable, off by default.
  static int cluster_ctas = -1;                           // largest cluster of 1024-thread CTAs this device schedules
  if (cluster_ctas < 0) {
    cluster_ctas = 0;
    if (cudaFuncSetAttribute(octree_cast_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      for (int c = 16; c >= 8 && cluster_ctas == 0; c -= 8) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c); cfg.blockDim = dim3(1024);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, octree_cast_kernel<true>, &cfg) == cudaSuccess && n >= 1) cluster_ctas = c;
      }
    }
    (void)cudaGetLastError();
  }
  const int rays_per_warp = 4;
  if (cluster_ctas > 0 && p->K <= cluster_ctas * 32 * rays_per_warp && getenv("ROBIR_OCTREE_CLUSTER")) {
    int ctas = (p->K + 31) / 32;                            // one warp per ray while the rays fit, then up to 4 per warp
    ctas = ctas > cluster_ctas ? cluster_ctas : (ctas < 1 ? 1 : ctas);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(1024); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, octree_cast_kernel<true>, params));
    return 0;
  }
  int per_sm = 0;
  RB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, octree_cast_kernel<false>, 256, 0));
  RB_REQUIRE(per_sm >= 1, "octree_cast: kernel does not fit on an SM");
  long long want = ((long long)p->K * 32 + 255) / 256;   // one warp per ray
  int grid = (int)(want < (long long)per_sm * sm_count ? want : (long long)per_sm * sm_count);
  if (grid < 1) grid = 1;
  void* args[] = {&params};
  RB_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)octree_cast_kernel<false>, dim3(grid), dim3(256), args, 0,
                                            (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
