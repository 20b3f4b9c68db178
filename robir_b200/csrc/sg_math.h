// Spherical-Gaussian shading math for the RobIR hot path, host+device, templated on the scalar type so the same
// source gives the forward value (T = float) and exact first derivatives (T = Dual<N>, forward-mode AD).
// Reference behaviour restated (not copied) from model/sg_render.py: hemisphere_int :62-81, lambda_trick :84-104,
// render_with_sg :343-565 (single view, metallic=None), visibility sample directions :117-146 / :204-240.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

namespace robir {

constexpr float kTiny = 1e-6f;        // sg_render.py:6
constexpr float kMuCos = 32.7080f;    // :381
constexpr float kLambdaCos = 0.0315f;  // :382
constexpr float kAlphaCos = 31.7003f;  // :383
constexpr float kPi = 3.14159265358979323846f;

// ---------------------------------------------------------------------------------------------------------------
// forward-mode dual number
// ---------------------------------------------------------------------------------------------------------------
template <int N>
struct Dual {
  float v;
  float d[N];
  RB_HD Dual() {}
  RB_HD Dual(float x) : v(x) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = 0.f;
  }
  RB_HD static Dual seed(float x, int k) {
    Dual r(x);
    r.d[k] = 1.f;
    return r;
  }
};

RB_HD float val(float x) { return x; }
template <int N>
RB_HD float val(const Dual<N>& x) { return x.v; }

// derivative propagation that treats "no dependence" (tangent exactly 0) as 0 even when the local slope is inf/nan,
// like reverse-mode autograd, which never visits a path without gradient (e.g. sqrt'(0) * 0 at a detached input).
RB_HD float dmul(float slope, float tangent) { return tangent == 0.f ? 0.f : slope * tangent; }

#define RB_DUAL_UNARY(name, fv, dfdx)                 \
  template <int N>                                    \
  RB_HD Dual<N> name(const Dual<N>& a) {              \
    Dual<N> r;                                        \
    const float x = a.v;                              \
    const float f = (fv);                             \
    const float g = (dfdx);                           \
    r.v = f;                                          \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = dmul(g, a.d[i]); \
    return r;                                         \
  }

RB_HD float t_sqrt(float x) { return sqrtf(x); }
RB_HD float t_exp(float x) { return expf(x); }
RB_HD float t_exp2(float x) { return exp2f(x); }
RB_HD float t_sin(float x) { return sinf(x); }
RB_HD float t_cos(float x) { return cosf(x); }
RB_HD float t_acos(float x) { return acosf(x); }
RB_HD float t_abs(float x) { return fabsf(x); }
RB_DUAL_UNARY(t_sqrt, sqrtf(x), 0.5f / f)
RB_DUAL_UNARY(t_exp, expf(x), f)
RB_DUAL_UNARY(t_exp2, exp2f(x), f * 0.69314718055994530942f)
RB_DUAL_UNARY(t_sin, sinf(x), cosf(x))
RB_DUAL_UNARY(t_cos, cosf(x), -sinf(x))
RB_DUAL_UNARY(t_acos, acosf(x), -1.f / sqrtf(1.f - x * x))
RB_DUAL_UNARY(t_abs, fabsf(x), (x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f)))

template <int N>
RB_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N>
RB_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N>
RB_HD Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N>
RB_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = dmul(b.v, a.d[i]) + dmul(a.v, b.d[i]);
  return r;
}
template <int N>
RB_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  const float inv = 1.f / b.v;
  r.v = a.v / b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = dmul(inv, a.d[i]) - dmul(r.v * inv, b.d[i]);
  return r;
}
template <int N>
RB_HD Dual<N> operator+(const Dual<N>& a, float b) { Dual<N> r = a; r.v = a.v + b; return r; }
template <int N>
RB_HD Dual<N> operator+(float b, const Dual<N>& a) { Dual<N> r = a; r.v = a.v + b; return r; }
template <int N>
RB_HD Dual<N> operator-(const Dual<N>& a, float b) { Dual<N> r = a; r.v = a.v - b; return r; }
template <int N>
RB_HD Dual<N> operator-(float b, const Dual<N>& a) { Dual<N> r = -a; r.v = b - a.v; return r; }
template <int N>
RB_HD Dual<N> operator*(const Dual<N>& a, float b) {
  Dual<N> r;
  r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = dmul(b, a.d[i]);
  return r;
}
template <int N>
RB_HD Dual<N> operator*(float b, const Dual<N>& a) { return a * b; }
template <int N>
RB_HD Dual<N> operator/(const Dual<N>& a, float b) { return a * (1.f / b); }
template <int N>
RB_HD Dual<N> operator/(float a, const Dual<N>& b) { return Dual<N>(a) / b; }

// torch.clamp(x, min=lo) / (max=hi) / torch.min(a, b): sub-gradient follows torch autograd
// (clamp passes the gradient where lo <= x <= hi; min/max of two tensors split ties evenly -- ties are measure-zero here).
RB_HD float t_clamp_min(float x, float lo) { return x < lo ? lo : x; }
RB_HD float t_clamp_max(float x, float hi) { return x > hi ? hi : x; }
template <int N>
RB_HD Dual<N> t_clamp_min(const Dual<N>& x, float lo) { return x.v < lo ? Dual<N>(lo) : x; }
template <int N>
RB_HD Dual<N> t_clamp_max(const Dual<N>& x, float hi) { return x.v > hi ? Dual<N>(hi) : x; }
RB_HD float t_min(float a, float b) { return a < b ? a : b; }
template <int N>
RB_HD Dual<N> t_min(const Dual<N>& a, const Dual<N>& b) {
  if (a.v < b.v) return a;
  if (b.v < a.v) return b;
  return (a + b) * 0.5f;
}

template <typename T>
struct V3 {
  T x, y, z;
};
template <typename T>
RB_HD T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T>
RB_HD V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T>
RB_HD V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T>
RB_HD V3<T> scale(const V3<T>& a, const T& s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T>
RB_HD V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// x / (||x|| + 1e-6)   (norm_axis, sg_render.py:107-108)
template <typename T>
RB_HD V3<T> norm_axis(const V3<T>& a) {
  T n = t_sqrt(dot(a, a)) + kTiny;
  return {a.x / n, a.y / n, a.z / n};
}
template <typename T>
RB_HD V3<T> lift3(const V3<float>& a) { return {T(a.x), T(a.y), T(a.z)}; }

// ---------------------------------------------------------------------------------------------------------------
// hemisphere_int (sg_render.py:62-81)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
RB_HD T hemisphere_int(T lam, T cos_beta) {
  lam = lam + kTiny;
  T inv = 1.f / lam;
  T t = t_sqrt(lam) * (1.6988f + 10.8438f * inv) / (1.f + 6.2201f * inv + 10.2415f * inv * inv);
  T inv_a = t_exp(-t);
  T s;
  if (val(cos_beta) >= 0.f) {
    T inv_b = t_exp(-t * t_clamp_min(cos_beta, 0.f));
    s = (1.f - inv_a * inv_b) / (1.f - inv_a + inv_b - inv_a * inv_b);
  } else {
    T b = t_exp(t * t_clamp_max(cos_beta, 0.f));
    s = (b - inv_a) / ((1.f - inv_a) * (b + 1.f));
  }
  T e1 = t_exp(-lam);
  T e2 = t_exp(-2.f * lam);
  T a_b = 2.f * kPi / lam * (e1 - e2);
  T a_u = 2.f * kPi / lam * (1.f - e1);
  return a_b * (1.f - s) + a_u * s;
}

// ---------------------------------------------------------------------------------------------------------------
// lambda_trick (sg_render.py:84-104): product of two SGs, lambda1 << lambda2.  mu handled by the caller as a scalar
// factor exp(diff) (mu3 = mu1*mu2*factor) so that RGB amplitudes stay outside the dual arithmetic where possible.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct SGProduct {
  V3<T> lobe;
  T lam;
  T factor;
};
template <typename T>
RB_HD SGProduct<T> lambda_trick(const V3<T>& lobe1_in, const T& lam1, const V3<T>& lobe2_in, const T& lam2) {
  T ratio = lam1 / lam2;
  V3<T> l1 = norm_axis(lobe1_in);
  V3<T> l2 = norm_axis(lobe2_in);
  T d = dot(l1, l2);
  T tmp = t_sqrt(ratio * ratio + 1.f + 2.f * ratio * d);
  tmp = t_min(tmp, ratio + 1.f);
  SGProduct<T> r;
  r.lam = lam2 * tmp;
  T a = ratio / tmp, b = 1.f / tmp;
  r.lobe = scale(l1, a) + scale(l2, b);
  r.factor = t_exp(lam2 * (tmp - ratio - 1.f));
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// per-point quantities of the specular branch that do not depend on the light lobe (sg_render.py:414-458)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct SpecPoint {
  V3<T> warp_lobe;   // warpBrdfSGLobes
  T warp_lam;        // warpBrdfSGLambdas
  T mu_base;         // (2/r^4)/pi * G/(4 d1 d2 + tiny); multiply by Fresnel per channel
  T v_dot_h;         // clamped
};
template <typename T>
RB_HD SpecPoint<T> spec_point(const V3<T>& n, const V3<T>& v, const T& rough) {
  SpecPoint<T> s;
  T inv_r4 = 2.f / (rough * rough * rough * rough);
  T vdl = t_clamp_min(dot(n, v), 0.f);
  V3<T> wl = scale(n, 2.f * vdl) - v;
  T wn = t_sqrt(dot(wl, wl)) + kTiny;
  wl = {wl.x / wn, wl.y / wn, wl.z / wn};
  s.warp_lobe = wl;
  s.warp_lam = inv_r4 / (4.f * vdl + kTiny);
  V3<T> h = wl + v;
  T hn = t_sqrt(dot(h, h)) + kTiny;
  h = {h.x / hn, h.y / hn, h.z / hn};
  s.v_dot_h = t_clamp_min(dot(v, h), 0.f);
  T d1 = t_clamp_min(dot(wl, n), 0.f);
  T d2 = t_clamp_min(dot(v, n), 0.f);
  T k = (rough + 1.f) * (rough + 1.f) / 8.f;
  T g1 = d1 / (d1 * (1.f - k) + k + kTiny);
  T g2 = d2 / (d2 * (1.f - k) + k + kTiny);
  s.mu_base = (inv_r4 / kPi) * (g1 * g2) / (4.f * d1 * d2 + kTiny);
  return s;
}
// Schlick-SG Fresnel (sg_render.py:438-439)
template <typename T>
RB_HD T fresnel(const T& spec_refl, const T& v_dot_h) {
  return spec_refl + (1.f - spec_refl) * t_exp2(-(5.55473f * v_dot_h + 6.8316f) * v_dot_h);
}

// One light lobe's specular contribution per unit (mu_light_c * brdf_vis * F_c): returns the scalar
//   K = mu_prime_factor * H(lambda', n.l') - alpha_cos * H(lambda_f, n.l_f)     scaled by factor_f * mu_base
// so that  spec_c = mu_light_c * brdf_vis * F_c * K          (sg_render.py:478-493)
template <typename T>
RB_HD T spec_lobe_kernel(const V3<T>& n, const SpecPoint<T>& sp, const V3<T>& lgt_lobe, const T& lgt_lam) {
  SGProduct<T> f = lambda_trick(lgt_lobe, lgt_lam, sp.warp_lobe, sp.warp_lam);
  SGProduct<T> p = lambda_trick(n, T(kLambdaCos), f.lobe, f.lam);
  T da = dot(p.lobe, n);
  T db = dot(f.lobe, n);
  T inner = kMuCos * p.factor * hemisphere_int(p.lam, da) - kAlphaCos * hemisphere_int(f.lam, db);
  return sp.mu_base * f.factor * inner;
}
// One light lobe's diffuse contribution per unit (mu_light_c * light_vis * albedo_c/pi)   (sg_render.py:511-529)
template <typename T>
RB_HD T diffuse_lobe_kernel(const V3<T>& n, const V3<T>& lgt_lobe, const T& lgt_lam) {
  SGProduct<T> p = lambda_trick(n, T(kLambdaCos), lgt_lobe, lgt_lam);
  T da = dot(p.lobe, n);
  T db = dot(lgt_lobe, n);
  return kMuCos * p.factor * hemisphere_int(p.lam, da) - kAlphaCos * hemisphere_int(lgt_lam, db);
}

// raw [7] SG -> normalised lobe, |lambda|, |mu|   (sg_render.py:364-366)
template <typename T>
struct LightSG {
  V3<T> lobe;
  T lam;
  T mu[3];
};
template <typename T>
RB_HD LightSG<T> decode_light(const T* raw) {
  LightSG<T> l;
  V3<T> a = {raw[0], raw[1], raw[2]};
  l.lobe = norm_axis(a);
  l.lam = t_abs(raw[3]);
  l.mu[0] = t_abs(raw[4]);
  l.mu[1] = t_abs(raw[5]);
  l.mu[2] = t_abs(raw[6]);
  return l;
}

// ---------------------------------------------------------------------------------------------------------------
// visibility sample direction around an axis (sg_render.py:123-146 diffuse, :204-240 specular)
//   frame axis a (diffuse: norm_axis(light lobe); specular: reflection dir, used un-normalised as in the reference),
//   phi_range = acos(1 - 0.95*sg_range/sharp),  dir = U cos(th) sin(ph) + V sin(th) sin(ph) + a cos(ph)
//   weight = exp(lam_w * (dir . axis_w - 1))
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
RB_HD void sample_dir(const V3<T>& a, const V3<T>& axis_w, const T& sharp, const T& lam_w, const T& sg_range,
                      float u_theta, float u_phi, V3<T>* dir, T* weight) {
  V3<T> z = {T(0.f), T(0.f), T(1.f)};
  V3<T> U = norm_axis(cross(z, a));
  V3<T> V = norm_axis(cross(a, U));
  T phi_range = t_acos((-0.95f * sg_range) / sharp + 1.f);
  float th = u_theta * 2.f * kPi;
  T ph = phi_range * u_phi;
  T sp = t_sin(ph), cp = t_cos(ph);
  float ct = cosf(th), st = sinf(th);
  V3<T> d = scale(U, sp * ct) + scale(V, sp * st) + scale(a, cp);
  *dir = d;
  *weight = t_exp(lam_w * (dot(d, axis_w) - 1.f));
}

}  // namespace robir
