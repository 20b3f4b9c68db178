// TEST INFRASTRUCTURE: host build (g++) of the host+device math headers used by the CUDA kernels
// (robir_b200/csrc/sg_math.h, octree_walk.h, loss_math.h) so their logic can be checked against the oracle on machines without a
// GPU.  Never linked into or reachable from the product library.
#include <cstring>
#include <vector>

#include "loss_math.h"
#include "octree_walk.h"
#include "sg_math.h"

using namespace robir;

extern "C" {

// lock-step octree cast, mirroring octree_cast_kernel (trace.cu) with the scalar helpers of octree_walk.h
void hc_octree_cast(const float* nodes, int n_nodes, const int* grid, int gx, int gy, int gz, const float* root,
                    const float* sdf_grad, const float* rays_o, const float* rays_d, int K, int o_div, int max_iter,
                    float refine_limit, float last_sdf, float* out_t, unsigned char* out_hit, float* out_x,
                    int* out_iters) {
  OctreeView o;
  o.nodes = reinterpret_cast<const OctNode*>(nodes);
  o.grid = grid; o.gx = gx; o.gy = gy; o.gz = gz; o.n_nodes = n_nodes;
  o.rminx = root[0]; o.rminy = root[1]; o.rminz = root[2]; o.rsizex = root[3]; o.rsizey = root[4]; o.rsizez = root[5];
  const bool secondary = max_iter > 0;
  const float eps = 1e-3f;
  std::vector<RayState> st(K);
  std::vector<float> org(3 * K);
  unsigned visits = 0, samples = 0;
  long long live = 0;
  for (int r = 0; r < K; ++r) {
    const float* d = rays_d + 3 * r;
    const float* po = rays_o + 3 * (r / o_div);
    float ox = po[0], oy = po[1], oz = po[2];
    if (secondary) { ox = ox + d[0] * 0.005f; oy = oy + d[1] * 0.005f; oz = oz + d[2] * 0.005f; }
    org[3 * r] = ox; org[3 * r + 1] = oy; org[3 * r + 2] = oz;
    ray_init(o, ox, oy, oz, d[0], d[1], d[2], eps, &st[r], &visits);
    live += st[r].live;
  }
  int it = 0;
  for (;; ++it) {
    if (live == 0) break;
    if (secondary && it > max_iter) break;
    float step = 0.001f;
    if (secondary) step = K > 100000 ? 0.01f : 0.005f;
    long long q = (long long)K * 10;
    q = q < 1 ? 1 : (q > 2000000 ? 2000000 : q);
    int ms = (int)(q / live);
    ms = ms < 1 ? 1 : (ms > 100 ? 100 : ms);
    long long next = 0;
    for (int r = 0; r < K; ++r) {
      if (!st[r].live) continue;
      const float* d = rays_d + 3 * r;
      ray_step(o, org[3 * r], org[3 * r + 1], org[3 * r + 2], d[0], d[1], d[2], eps, ms, step, last_sdf, &st[r],
               &visits, &samples);
      next += st[r].live;
    }
    live = next;
  }
  *out_iters = it;
  for (int r = 0; r < K; ++r) {
    const float* d = rays_d + 3 * r;
    float t = st[r].t;
    if (st[r].ptr >= 0) {
      const float* g = sdf_grad + 3 * (size_t)st[r].ptr;
      t = refine_t(st[r], g[0], g[1], g[2], d[0], d[1], d[2], refine_limit);
    }
    const float* po = rays_o + 3 * (r / o_div);
    out_t[r] = t;
    out_hit[r] = st[r].ptr >= 0;
    for (int c = 0; c < 3; ++c) out_x[3 * r + c] = t * d[c] + po[c];
  }
}

// SG render forward + backward for n points, mirroring sg_render_{fwd,bwd}_kernel (sg.cu)
void hc_sg_render(int n, int M, int Mi, const float* normal, const float* view, const float* rough,
                  const float* albedo, float spec_refl, const float* lgt, const float* ind_lgt,
                  const float* light_vis, const float* bv_dir, const float* bv_ind, const float* ind_integral,
                  float* out /*[n][7][3]: rgb spec diff shadow irgb ispec idiff*/,
                  const float* g_out /*[n][7][3] upstream*/, float* g_lgt /*[M][7]*/, float* g_ind_lgt,
                  float* g_light_vis, float* g_bv_dir, float* g_bv_ind, float* g_rough, float* g_albedo,
                  float* g_spec_refl, float* g_ind_integral, float* g_normal /*[n][3] or null*/) {
  // the widest instantiation of sg_render_bwd_kernel (NG = true): specular duals carry the normal in slots 10-12,
  // diffuse duals in slots 8-10
  typedef Dual<13> DS;
  typedef Dual<11> DD;
  memset(g_lgt, 0, sizeof(float) * M * 7);
  *g_spec_refl = 0.f;
  for (int i = 0; i < n; ++i) {
    V3<float> nrm = {normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]};
    V3<float> vw = {view[3 * i], view[3 * i + 1], view[3 * i + 2]};
    const float* alb = albedo + 3 * i;
    SpecPoint<float> sp = spec_point<float>(nrm, vw, rough[i]);
    float F = fresnel<float>(spec_refl, sp.v_dot_h);
    float v[15] = {0};
    for (int m = 0; m < M; ++m) {
      LightSG<float> l = decode_light<float>(lgt + 7 * m);
      float lv = light_vis[i * M + m];
      float Ks = spec_lobe_kernel<float>(nrm, sp, l.lobe, l.lam);
      float Kd = diffuse_lobe_kernel<float>(nrm, l.lobe, l.lam);
      for (int c = 0; c < 3; ++c) {
        v[c] += l.mu[c] * bv_dir[i] * F * Ks;
        v[3 + c] += l.mu[c] * lv * (alb[c] / kPi) * Kd;
        v[6 + c] += lv * l.mu[c];
        v[9 + c] += l.mu[c];
      }
    }
    for (int m = 0; m < Mi; ++m) {
      LightSG<float> l = decode_light<float>(ind_lgt + (i * Mi + m) * 7);
      float Ks = spec_lobe_kernel<float>(nrm, sp, l.lobe, l.lam);
      for (int c = 0; c < 3; ++c) v[12 + c] += l.mu[c] * bv_ind[i] * F * Ks;
    }
    float* o = out + i * 21;
    float gs[3], gd[3], gis[3], gid[3];
    for (int c = 0; c < 3; ++c) {
      float spec = fmaxf(v[c], 0.f), diff = fmaxf(v[3 + c], 0.f);
      o[c] = spec + diff; o[3 + c] = spec; o[6 + c] = diff; o[9 + c] = v[6 + c] / fmaxf(v[9 + c], 1e-4f);
      float ispec = fmaxf(v[12 + c], 0.f), idiff = ind_integral[3 * i + c] * (alb[c] / kPi);
      o[12 + c] = ispec + idiff; o[15 + c] = ispec; o[18 + c] = idiff;
      const float* g = g_out + i * 21;
      gs[c] = v[c] >= 0.f ? g[c] + g[3 + c] : 0.f;
      gd[c] = v[3 + c] >= 0.f ? g[c] + g[6 + c] : 0.f;
      gis[c] = v[12 + c] >= 0.f ? g[12 + c] + g[15 + c] : 0.f;
      gid[c] = g[12 + c] + g[18 + c];
    }
    // backward
    V3<DS> nS = {DS::seed(nrm.x, 10), DS::seed(nrm.y, 11), DS::seed(nrm.z, 12)}, vS = lift3<DS>(vw);
    float gn[3] = {0.f, 0.f, 0.f};
    SpecPoint<DS> spd = spec_point<DS>(nS, vS, DS::seed(rough[i], 7));
    DS Fd = fresnel<DS>(DS::seed(spec_refl, 8), spd.v_dot_h);
    V3<DD> nD = {DD::seed(nrm.x, 8), DD::seed(nrm.y, 9), DD::seed(nrm.z, 10)};
    float acc[7] = {0};
    auto spec_lobe = [&](const float* raw, float bv, const float* g, float* g_raw, int slot) {
      DS r[7];
      for (int k = 0; k < 7; ++k) r[k] = DS::seed(raw[k], k);
      LightSG<DS> l = decode_light<DS>(r);
      DS Ks = spec_lobe_kernel<DS>(nS, spd, l.lobe, l.lam);
      DS common = DS::seed(bv, 9) * Fd * Ks;
      DS tot(0.f);
      for (int c = 0; c < 3; ++c) tot = tot + (l.mu[c] * common) * g[c];
      for (int k = 0; k < 7; ++k) g_raw[k] = tot.d[k];
      acc[0] += tot.d[7]; acc[1] += tot.d[8]; acc[slot] += tot.d[9];
      for (int k = 0; k < 3; ++k) gn[k] += tot.d[10 + k];
    };
    for (int m = 0; m < M; ++m) {
      float g_raw[7];
      spec_lobe(lgt + 7 * m, bv_dir[i], gs, g_raw, 2);
      DD r[7];
      for (int k = 0; k < 7; ++k) r[k] = DD::seed(lgt[7 * m + k], k);
      LightSG<DD> l = decode_light<DD>(r);
      DD lv = DD::seed(light_vis[i * M + m], 7);
      DD Kd = diffuse_lobe_kernel<DD>(nD, l.lobe, l.lam);
      DD tot(0.f);
      for (int c = 0; c < 3; ++c) {
        DD base = l.mu[c] * lv * Kd;
        tot = tot + base * ((alb[c] / kPi) * gd[c]);
        acc[4 + c] += base.v * gd[c] / kPi;
      }
      for (int k = 0; k < 7; ++k) g_lgt[7 * m + k] += g_raw[k] + tot.d[k];
      g_light_vis[i * M + m] = tot.d[7];
      for (int k = 0; k < 3; ++k) gn[k] += tot.d[8 + k];
    }
    for (int m = 0; m < Mi; ++m) spec_lobe(ind_lgt + (i * Mi + m) * 7, bv_ind[i], gis, g_ind_lgt + (i * Mi + m) * 7, 3);
    if (g_normal) for (int k = 0; k < 3; ++k) g_normal[3 * i + k] = gn[k];
    g_rough[i] = acc[0];
    *g_spec_refl += acc[1];
    g_bv_dir[i] = acc[2];
    g_bv_ind[i] = acc[3];
    for (int c = 0; c < 3; ++c) {
      g_albedo[3 * i + c] = acc[4 + c] + gid[c] * ind_integral[3 * i + c] / kPi;
      g_ind_integral[3 * i + c] = gid[c] * (alb[c] / kPi);
    }
  }
}

// sample directions forward + backward (mirrors sample_dirs_{fwd,bwd}_kernel, vis.cu)
void hc_sample_dirs(int K, int S, const float* axis_f, const float* axis_w, const float* sharp, const float* lam_w,
                    float sg_range, const float* u_theta, const float* u_phi, int renorm, float* dirs, float* w,
                    const float* g_dirs, const float* g_w, float* g_axis_f, float* g_axis_w, float* g_sharp,
                    float* g_lam_w, float* g_sg_range) {
  typedef Dual<9> D;
  *g_sg_range = 0.f;
  for (int k = 0; k < K; ++k) {
    float g[9] = {0};
    for (int s = 0; s < S; ++s) {
      const int idx = k * S + s;
      V3<float> af = {axis_f[3 * k], axis_f[3 * k + 1], axis_f[3 * k + 2]};
      V3<float> aw = {axis_w[3 * k], axis_w[3 * k + 1], axis_w[3 * k + 2]};
      if (renorm) { af = norm_axis(af); aw = af; }
      V3<float> d; float ww;
      sample_dir<float>(af, aw, sharp[k], lam_w[k], sg_range, u_theta[idx], u_phi[idx], &d, &ww);
      dirs[3 * idx] = d.x; dirs[3 * idx + 1] = d.y; dirs[3 * idx + 2] = d.z; w[idx] = ww;
      V3<D> afd = {D::seed(axis_f[3 * k], 0), D::seed(axis_f[3 * k + 1], 1), D::seed(axis_f[3 * k + 2], 2)};
      V3<D> awd;
      if (renorm) { afd = norm_axis(afd); awd = afd; }
      else awd = {D::seed(axis_w[3 * k], 3), D::seed(axis_w[3 * k + 1], 4), D::seed(axis_w[3 * k + 2], 5)};
      V3<D> dd; D wd;
      sample_dir<D>(afd, awd, D::seed(sharp[k], 6), D::seed(lam_w[k], 7), D::seed(sg_range, 8), u_theta[idx],
                    u_phi[idx], &dd, &wd);
      for (int i = 0; i < 9; ++i)
        g[i] += g_dirs[3 * idx] * dd.x.d[i] + g_dirs[3 * idx + 1] * dd.y.d[i] + g_dirs[3 * idx + 2] * dd.z.d[i] +
                g_w[idx] * wd.d[i];
    }
    for (int i = 0; i < 3; ++i) { g_axis_f[3 * k + i] = g[i]; g_axis_w[3 * k + i] = g[3 + i]; }
    g_sharp[k] = g[6]; g_lam_w[k] = g[7]; *g_sg_range += g[8];
  }
}


// fused PBR loss, mirroring pbr_loss_kernel (loss.cu) sequentially with the helpers of loss_math.h.
// losses [5] = total, rgb, kl, smooth, white; gradients as in robir_loss_params (contiguous inputs, no compaction).
void hc_pbr_loss(int N, int n_lat, int M, int l2, const float* sg_rgb, const float* indir_rgb, const float* gt,
                 const unsigned char* mask, float adapt, const float* albedo, const float* albedo_r, const float* rough,
                 const float* rough_r, const float* z, const unsigned char* z_valid, const float* lgt, float w_rgb,
                 float w_kl, float w_smooth, float rho, float* losses, float* g_pred, float* g_adapt, float* g_albedo,
                 float* g_albedo_r, float* g_rough, float* g_rough_r, float* g_z, float* g_lgt) {
  bool live;
  const float shift = loss_shift(adapt, &live);
  const float inv_s02 = powf(shift, -0.2f), inv_N = 1.f / (float)N;
  double rgb = 0, dshift = 0, sa = 0, sr = 0, white = 0;
  for (int i = 0; i < N; ++i) {
    const float m = mask[i] ? 1.f : 0.f;
    for (int c = 0; c < 3; ++c) {
      float per, d_pred, d_shift;
      loss_rgb_channel(sg_rgb[3 * i + c] + indir_rgb[3 * i + c], gt[3 * i + c], inv_s02, l2, m, &per, &d_pred, &d_shift);
      rgb += per;
      dshift += d_shift;
      g_pred[3 * i + c] = w_rgb * inv_N * d_pred;
      const float da = albedo[3 * i + c] - albedo_r[3 * i + c];
      sa += fabsf(da);
      g_albedo[3 * i + c] = w_smooth * loss_sgn(da) * inv_N * (1.f / 3.f);
      g_albedo_r[3 * i + c] = -g_albedo[3 * i + c];
    }
    const float dr = rough[i] - rough_r[i];
    sr += fabsf(dr);
    g_rough[i] = w_smooth * 0.2f * loss_sgn(dr) * inv_N;
    g_rough_r[i] = -g_rough[i];
  }
  double col[32] = {0};
  double nv = 0;
  for (int i = 0; i < n_lat; ++i) {
    if (z_valid && !z_valid[i]) continue;
    nv += 1;
    for (int j = 0; j < 32; ++j) col[j] += loss_sigmoid(z[32 * i + j]);
  }
  const float n_valid = nv > 1 ? (float)nv : 1.f;
  float dkl[32];
  double kl = 0;
  for (int j = 0; j < 32; ++j) {
    float d;
    kl += loss_kl_column(rho, (float)(col[j] / n_valid), &d);
    dkl[j] = d * (1.f / 32.f);
  }
  kl /= 32.0;
  for (int i = 0; i < n_lat; ++i)
    for (int j = 0; j < 32; ++j) {
      const float s = loss_sigmoid(z[32 * i + j]);
      g_z[32 * i + j] = (z_valid && !z_valid[i]) ? 0.f : w_kl / n_valid * dkl[j] * s * (1.f - s);
    }
  for (int i = 0; i < M; ++i) {
    const float x[3] = {lgt[7 * i + 4], lgt[7 * i + 5], lgt[7 * i + 6]};
    float gx[3];
    white += loss_white_lobe(x, gx);
    for (int k = 0; k < 4; ++k) g_lgt[7 * i + k] = 0.f;
    for (int j = 0; j < 3; ++j) g_lgt[7 * i + 4 + j] = 0.01f / (float)M * gx[j];
  }
  const float rgb_l = (float)rgb * inv_N;
  const float smooth = (float)sa * inv_N * (1.f / 3.f) + 0.2f * (float)sr * inv_N;
  const float white_l = 0.01f * (float)white / (float)M;
  losses[0] = w_rgb * rgb_l + w_kl * (float)kl + w_smooth * smooth + white_l;
  losses[1] = rgb_l; losses[2] = (float)kl; losses[3] = smooth; losses[4] = white_l;
  g_adapt[0] = live ? w_rgb * inv_N * (float)dshift * (-0.2f) * powf(shift, -1.2f) * 10.f : 0.f;
}

}  // extern "C"
