// Per-ray octree walk over the cached-SDF octree, host+device.  Behaviour restated from the reference
// utils/octree.py: intersect_box :41-57, inside_box :19-29, which_oct_cell :32-38, Octree.query :217-265,
// OctreeSDF.fast_volume_render :459-471, multi_step_cast :493-585, cast :421-438 (quirks: SURVEY.md A.3).
//
// Packed node record (32 B, one DRAM sector per visit):  {min.xyz, size.x} {size.yz, child_base(int bits), sdf_val}
//   child_base = links[node][0] for internal nodes (children are allocated as 8 consecutive records, octree.py:163-169),
//   -1 for leaves.  hit_ptr = (max(sdf_val,0) <= 1e-4) is derived from sdf_val (octree.py:406-408).
// This translation unit must be compiled WITHOUT floating-point contraction (nvcc -fmad=false / g++ -ffp-contract=off):
// the reference evaluates every product and sum as a separately rounded fp32 op.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

namespace robir {

struct OctNode {
  float minx, miny, minz, sizex;
  float sizey, sizez;
  int child_base;
  float sdf_val;
};

struct OctreeView {
  const OctNode* nodes;   // [n_nodes]
  const int* grid;        // [gx*gy*gz] base-grid cell -> node
  int gx, gy, gz;
  int n_nodes;
  float rminx, rminy, rminz, rsizex, rsizey, rsizez;  // root box
};

// torch.minimum / torch.maximum propagate NaN (fminf/fmaxf do not)
RB_HD float tmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
RB_HD float tmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

RB_HD bool inside_root_open(const OctreeView& o, float x, float y, float z) {
  const float rx = (x - o.rminx) / o.rsizex, ry = (y - o.rminy) / o.rsizey, rz = (z - o.rminz) / o.rsizez;
  return (rx < 1.f) && (ry < 1.f) && (rz < 1.f) && (rx > 0.f) && (ry > 0.f) && (rz > 0.f);
}

RB_HD int clamp01_trunc(float v) {
  // (v).long() then clip(0,1); NaN/huge never reach here for points inside the root box
  int i = (int)v;
  return i < 0 ? 0 : (i > 1 ? 1 : i);
}

#if defined(__CUDA_ARCH__)
#define RB_LDG_NODE(dst, ptr)                                             \
  {                                                                       \
    const float4* p4 = reinterpret_cast<const float4*>(ptr);              \
    float4 a = __ldg(p4), b = __ldg(p4 + 1);                              \
    dst.minx = a.x; dst.miny = a.y; dst.minz = a.z; dst.sizex = a.w;      \
    dst.sizey = b.x; dst.sizez = b.y; dst.child_base = __float_as_int(b.z); dst.sdf_val = b.w; \
  }
#else
#define RB_LDG_NODE(dst, ptr) dst = *(ptr);
#endif

// Octree.query for one point; returns -1 outside the (open) root box.  *visits counts node records read.
RB_HD int oct_query(const OctreeView& o, float x, float y, float z, OctNode* out, unsigned* visits) {
  if (!inside_root_open(o, x, y, z)) return -1;
  int ix = (int)floorf(((x - o.rminx) / o.rsizex) * (float)o.gx);
  int iy = (int)floorf(((y - o.rminy) / o.rsizey) * (float)o.gy);
  int iz = (int)floorf(((z - o.rminz) / o.rsizez) * (float)o.gz);
  // strictly-inside points can still round to the upper index; the reference would raise an index error there.
  ix = ix >= o.gx ? o.gx - 1 : ix; iy = iy >= o.gy ? o.gy - 1 : iy; iz = iz >= o.gz ? o.gz - 1 : iz;
  int ptr = o.grid[(ix * o.gy + iy) * o.gz + iz];
  OctNode nd;
  RB_LDG_NODE(nd, o.nodes + ptr);
  ++*visits;
  while (nd.child_base >= 0) {
    const int cx = clamp01_trunc(((x - nd.minx) / nd.sizex) * 2.f);
    const int cy = clamp01_trunc(((y - nd.miny) / nd.sizey) * 2.f);
    const int cz = clamp01_trunc(((z - nd.minz) / nd.sizez) * 2.f);
    ptr = nd.child_base + 4 * cx + 2 * cy + cz;
    RB_LDG_NODE(nd, o.nodes + ptr);
    ++*visits;
  }
  *out = nd;
  return ptr;
}

// intersect_box(forward_only=True): returns far; near/valid through pointers (may be null)
RB_HD float box_far(float bminx, float bminy, float bminz, float bsx, float bsy, float bsz, float ox, float oy,
                    float oz, float dx, float dy, float dz, float* near_out, bool* valid_out) {
  const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
  const float ax = (bminx - ox) * ix, ay = (bminy - oy) * iy, az = (bminz - oz) * iz;
  const float bx = (bsx + bminx - ox) * ix, by = (bsy + bminy - oy) * iy, bz = (bsz + bminz - oz) * iz;
  const float near = tmax(tmax(tmin(ax, bx), tmin(ay, by)), tmin(az, bz));
  const float far = tmin(tmin(tmax(ax, bx), tmax(ay, by)), tmax(az, bz));
  if (near_out) *near_out = tmax(near, 0.f);
  if (valid_out) *valid_out = (near <= far) && (far >= 0.f);
  return far;
}

// torch.linspace(0, 1, steps)[p] in fp32 (ATen: symmetric evaluation around the midpoint)
RB_HD float linspace01(int p, int steps) {
  const float step = 1.0f / (float)(steps - 1);
  return (p < steps / 2) ? (0.0f + step * (float)p) : (1.0f - step * (float)(steps - p - 1));
}

// fast_volume_render for one ray: samples at t[1..n], returns t[first] with t[p] = linspace01(p,n+1)*n*step + step
RB_HD float micro_march(const OctreeView& o, float px, float py, float pz, float dx, float dy, float dz, int n_samp,
                        float step, float last_node_sdf, unsigned* visits, unsigned* samples) {
  int first = n_samp;
  for (int c = 0; c < n_samp; ++c) {
    const float t = linspace01(c + 1, n_samp + 1) * (float)n_samp * step + step;
    const float x = px + dx * t, y = py + dy * t, z = pz + dz * t;
    OctNode nd;
    const int ptr = oct_query(o, x, y, z, &nd, visits);
    ++*samples;
    const float sdf = ptr >= 0 ? nd.sdf_val : last_node_sdf;  // sdf_val[-1] reads the last node (A.3)
    if (sdf <= step) { first = c; break; }
  }
  return linspace01(first, n_samp + 1) * (float)n_samp * step + step;
}

// Per-ray state carried across the lock-step iterations of multi_step_cast
struct RayState {
  float t;        // accumulated distance along the (possibly biased) ray
  float px, py, pz;
  int ptr;        // current node, -1 = left the box / never entered
  bool live;      // 'k' in the reference
  OctNode node;   // record of ptr (valid when ptr >= 0)
};

RB_HD bool node_is_hit(const OctNode& nd) { return tmax(nd.sdf_val, 0.f) <= 1e-4f; }

RB_HD void ray_init(const OctreeView& o, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                    RayState* s, unsigned* visits) {
  float near;
  bool valid;
  box_far(o.rminx, o.rminy, o.rminz, o.rsizex, o.rsizey, o.rsizez, ox, oy, oz, dx, dy, dz, &near, &valid);
  s->t = valid ? near + eps : -1.f;
  s->ptr = -1;
  s->px = s->py = s->pz = 0.f;
  if (valid) {
    s->px = ox + s->t * dx; s->py = oy + s->t * dy; s->pz = oz + s->t * dz;
    s->ptr = oct_query(o, s->px, s->py, s->pz, &s->node, visits);
  }
  s->live = s->ptr >= 0;
}

// One lock-step iteration for one live ray (octree.py:528-573).  multi_samp/step are the batch-level values.
RB_HD void ray_step(const OctreeView& o, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                    int multi_samp, float step, float last_node_sdf, RayState* s, unsigned* visits,
                    unsigned* samples) {
  const OctNode& b = s->node;
  float far = box_far(b.minx, b.miny, b.minz, b.sizex, b.sizey, b.sizez, s->px, s->py, s->pz, dx, dy, dz, nullptr,
                      nullptr);
  if (far < (float)multi_samp * step)
    far = micro_march(o, s->px, s->py, s->pz, dx, dy, dz, multi_samp, step, last_node_sdf, visits, samples);
  s->t = s->t + (far + eps);
  s->px = ox + s->t * dx; s->py = oy + s->t * dy; s->pz = oz + s->t * dz;
  if (!inside_root_open(o, s->px, s->py, s->pz)) {
    s->live = false;
    s->ptr = -1;
    return;
  }
  s->ptr = oct_query(o, s->px, s->py, s->pz, &s->node, visits);
  s->live = (s->ptr >= 0) && !node_is_hit(s->node);
}

// cast()'s first-order plane refinement (octree.py:427-434); grad = unit sdf_grad[ptr]; lim = fp32(10 * min_step)
RB_HD float refine_t(const RayState& s, float gx, float gy, float gz, float dx, float dy, float dz, float lim) {
  const OctNode& b = s.node;
  const float cx = b.minx + b.sizex * 0.5f, cy = b.miny + b.sizey * 0.5f, cz = b.minz + b.sizez * 0.5f;
  const float fx = cx - gx * b.sdf_val, fy = cy - gy * b.sdf_val, fz = cz - gz * b.sdf_val;
  const float dist = ((fx - s.px) * gx + (fy - s.py) * gy) + (fz - s.pz) * gz;
  float speed = (dx * gx + dy * gy) + dz * gz;
  if (speed == 0.f) speed = 1e-4f;
  float dt = dist / speed;
  dt = dt < -lim ? -lim : (dt > lim ? lim : dt);  // torch.clamp keeps NaN
  return s.t + dt;
}

}  // namespace robir
