/* robir_b200 C ABI -- the drop-in boundary of the B200-native RobIR hot path.
 *
 * The reference (ingra14m/RobIR) is pure Python/PyTorch and has no FFI; its seams are Python attributes
 * (SURVEY.md section 8b).  This library is what a binding for those seams calls: plain device pointers, sizes and
 * scalars, a CUDA stream, an int status (0 = ok; otherwise robir_last_error() describes the failure).  No allocation,
 * no host synchronisation and no C++ exceptions cross the boundary; all buffers (including workspaces) are owned by
 * the caller.  Every pointer is a device pointer to contiguous fp32 data unless stated otherwise.
 * "file:line" citations are into the reference tree.
 */
#ifndef ROBIR_B200_H_
#define ROBIR_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plumbing ---------------------------------------------------------------------------------------------------- */
const char* robir_last_error(void);
int robir_abi_version(void);
int robir_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- weight packing (torch Linear weight [N][K] row-major) ------------------------------------------------------- */
/* Wt[Kpad][Npad] = scale * W[:, k_begin : k_begin+k_count]^T, zero padded. */
int robir_pack_transpose(const float* W, int N, int K, int k_begin, int k_count, float* Wt, int Kpad, int Npad,
                         float scale, void* stream);
/* out[Npad][Kpad] = W[:, k_begin : k_begin+k_count], zero padded. */
int robir_pack_window(const float* W, int N, int K, int k_begin, int k_count, float* out, int Npad, int Kpad,
                      void* stream);
/* nn.utils.weight_norm folding (model/neus_model.py:378-379): Wt[k][n] = g[n] v[n][k] / ||v[n]||, rows
 * [n_begin, n_begin+n_count) of v, transposed and zero padded to [Kpad][Npad]. */
int robir_pack_wn_transpose(const float* v, const float* g, int N, int K, int n_begin, int n_count, float* Wt,
                            int Kpad, int Npad, void* stream);
int robir_pack_wn_row(const float* v, const float* g, int K, int n, float* out, void* stream);

/* ---- a5: NeuS SDF network, value + normal + feature in one launch --------------------------------------------------
 * replaces ImplicitNetworkMy.forward / .gradient (model/neus_model.py:785-818) over SDFNetwork.forward (:385-417). */
typedef struct {
  const float* pts; int n;          /* [n][3] */
  float in_scale;                   /* 2.0: stage-2 points (normalize(), :785-786); 1.0: NeuS coordinates */
  float sdf_scale, feat_scale;      /* 0.5 / 0.5 for ImplicitNetworkMy.forward (:788-792) */
  const float* Wt[8];               /* folded, transposed layers 0..7: [64|256][256] */
  const float* bias[8];             /* [256] zero padded */
  const float* w8_sdf;              /* [256] folded row 0 of layer 8 */
  const float* b8;                  /* [257] */
  const float* Wt8_feat;            /* [256][256] folded rows 1..256 of layer 8, transposed (NULL if feat == NULL) */
  float* sdf;                       /* [n] */
  float* grad;                      /* [n][3] or NULL  (d sdf / d p, forward-mode) */
  float* feat;                      /* [n][256] or NULL */
  const int* n_active;              /* optional device scalar: only points [0, min(*n_active, n)) are evaluated (hit rays
                                       compacted to the front of a fixed-capacity batch); the other outputs are 0 */
} robir_sdf_params;
int robir_sdf_eval(const robir_sdf_params* p, int sm_count, void* stream);

/* The same network on the tensor cores (tcgen05 / TMEM, scaled fp16 hi/lo 3-term products = fp32 parity): persistent
 * kernel, 128-row tiles (32 points x value + 3 tangent rows, or 128 points without the normal), activations in TMEM,
 * Softplus(100) and its derivative applied by the epilogue (csrc/sdf_tc.cu).  img: 8 layer images (9 with features) of
 * robir_tc_pack_layer(Wt_l, 256, 256, 64 | 256, transpose = 1, n_halves = 2, terms = 3) built from the folded transposed
 * weights above, layer 4 pre-multiplied by 1/sqrt(2) (the skip concat's scale); bias [8][256].  Points at or beyond
 * *n_active are not evaluated (the caller zero-fills). */
typedef struct {
  const float* pts; int n;
  float in_scale, sdf_scale, feat_scale;
  const void* img;
  const float* bias;                /* [8][256] */
  const float* w8_sdf;              /* [256] */
  const float* b8;                  /* [257] */
  float* sdf; float* grad; float* feat;
  const int* n_active;
} robir_sdf_tc_params;
int robir_sdf_tc(const robir_sdf_tc_params* p, int sm_count, void* stream);

/* ---- a2: camera rays: rend_util.get_camera_params + lift (utils/rend_util.py:51-97), one 4x4 pose ----------------- */
int robir_camera_rays(int N, const float* uv /*[N][2]*/, const float* pose /*[4][4]*/, const float* K /*[3][3]*/,
                      float* dirs /*[N][3]*/, void* stream);

/* ---- a4: octree surface tracer: OctreeSDF.cast (utils/octree.py:421-438) over multi_step_cast (:493-585) and
 * fast_volume_render (:459-471), + OctreeTracing.forward (model/octree_tracing.py:43-60) --------------------------- */
typedef struct {
  const void* nodes;                /* [n_nodes] 32-byte records {min.xyz, size.xyz, child_base(int), sdf_val} */
  const int* grid;                  /* [gx][gy][gz] base cell -> node (Octree.cache_index, :143) */
  int gx, gy, gz, n_nodes;
  float rminx, rminy, rminz, rsizex, rsizey, rsizez;   /* root box (Octree.whole_box) */
} robir_octree_view;
typedef struct {
  robir_octree_view view;
  const float* sdf_grad;            /* [n_nodes][3] unit gradients at node centres (:392-397) */
  const float* rays_o;              /* [K / o_div][3] */
  const float* rays_d;              /* [K][3] */
  int K, o_div;                     /* ray r starts at origin r / o_div */
  int max_iter;                     /* -1 primary; 32 secondary (origin bias 0.005, iteration cap; :504-505, :526) */
  float eps;                        /* 1e-3 */
  float refine_limit;               /* fp32(10 * min_step) (:433) */
  float last_node_sdf;              /* sdf_val[-1]: what an out-of-box micro-march sample reads */
  float* state_t; int* state_ptr;   /* [K] workspaces */
  float* out_t; float* out_x;       /* [K], [K][3] */
  unsigned char* out_hit;           /* [K] */
  unsigned* counters;               /* [robir_octree_counters_len()] zero-initialised: live rays per lock-step
                                       iteration, then {node visits, micro-march samples, iterations} */
} robir_octree_cast_params;
int robir_octree_counters_len(void);
int robir_octree_cast(const robir_octree_cast_params* p, int sm_count, void* stream);

/* ---- a3: IDR sphere tracer: RayTracing.forward / sphere_tracing / ray_sampler / secant / minimal_sdf_points
 * (model/ray_tracing.py:26-326) with rend_util.get_sphere_intersection (utils/rend_util.py:141-163).  The SDF network
 * is evaluated inline by persistent march kernels; no host synchronisation.  f(p) = net(in_scale p)[0] * out_scale. */
typedef struct {
  const float* Wt[8];   /* folded weight-norm layers 0..7, packed [Kpad][256] (robir_pack_wn_transpose) */
  const float* bias[8];
  const float* w8_sdf;  /* [256] folded row 0 of layer 8 (robir_pack_wn_row) */
  const float* b8;
} robir_sdf_net;
typedef struct {
  robir_sdf_net net;
  int N, o_div;                      /* rays; rays per origin (cam_loc is [N / o_div][3]) */
  const float* cam_loc;
  const float* ray_dirs;             /* [N][3] */
  const unsigned char* object_mask;  /* [N] or NULL (all true) */
  float in_scale, out_scale;
  float radius, sdf_threshold, line_search_step;
  int line_step_iters, sphere_tracing_iters, n_steps, n_secant_steps, training;
  const float* uniform_steps;        /* [n_steps]: the uniform_(0,1) draw of minimal_sdf_points (:305), training only */
  float* points;                     /* out [N][3] */
  unsigned char* net_mask;           /* out [N] */
  float* dists;                      /* out [N] */
  /* caller-owned workspace */
  float *acc_s, *acc_e, *min_dis, *max_dis; /* [N] each */
  unsigned char* flags;              /* [N] */
  int *samp_list, *sec_list;         /* [N] each */
  float* sec_state;                  /* [N][4] */
  int* min_list;                     /* [N] */
  float* vals;                       /* [N * n_steps] */
  int* counters;                     /* [8] zero-initialised; on return: sampler rays, secant rays, min-sdf rays,
                                        loop flag, executed SDF queries */
} robir_sphere_trace_params;
int robir_sphere_trace(const robir_sphere_trace_params* p, int sm_count, void* stream);
int robir_sphere_trace_launches(int training);

/* ---- a10/a11: visibility sample directions (model/sg_render.py:123-146 and :204-240) ------------------------------ */
int robir_sample_dirs_fwd(int K, int S, const float* axis_f, const float* axis_w, const float* sharp,
                          const float* lam_w, const float* sg_range /*[1]*/, const float* u_theta /*[K][S]*/,
                          const float* u_phi, int renorm_axis, float* dirs /*[K*S][3]*/, float* w /*[K*S]*/,
                          void* stream);
int robir_sample_dirs_bwd(int K, int S, const float* axis_f, const float* axis_w, const float* sharp,
                          const float* lam_w, const float* sg_range, const float* u_theta, const float* u_phi,
                          int renorm_axis, const float* g_dirs, const float* g_w, float* g_axis_f, float* g_axis_w,
                          float* g_sharp, float* g_lam_w, float* g_sg_range /*[1], zero-initialised*/, void* stream);

/* ---- a12: VisNetwork layer 0, factorised: tab[n][256] = PE10(x[n][3]) . Wt[64][256] (+ bias) ---------------------- */
int robir_pe_linear(const float* x, int n, const float* Wt, const float* bias, float* tab, void* stream);

/* ---- a10: (point, direction) pair lists.  live(i,j) = n_i . dir_j > 1e-6 (model/sg_render.py:155, :246) ----------- */
/* tile_rows = 64 (FFMA engine) or 128 (tensor-core engine).  pad_points != 0: every point's rows are padded to a tile
 * multiple (one point per tile); 0: points are packed back to back, only the last tile is padded (rowB = -1). */
int robir_diffuse_rows(int n, int M, int S, int tile_rows, int pad_points, const float* normals, const float* dirs,
                       uint32_t* bits /*[n][M]*/, int* lobe_off /*[n][M+1]*/, int* start /*[n]*/, int* rowA,
                       int* rowB /*[n*roundup(M*S,tile_rows)]*/, int* n_tiles /*[1]*/,
                       long long* n_pairs /*[1], accumulated*/, void* stream);
/* a_mod > 0: the n "points" are copies of the same a_mod points (rowA = (q / S) % a_mod, normals [a_mod][3]) */
int robir_spec_rows(int n, int S, int rows_padded, int tile_rows, int a_mod, const float* normals, const float* dirs,
                    int* rowA, int* rowB, int* n_tiles, long long* n_pairs, void* stream);
/* sampling inputs shared by the direct and the indirect get_specular_visibility call (model/sg_render.py:198-225 on the
 * warp of :417-428), written twice (rows i and n + i): ref / wl [2n][3], sharp [2n], sg_range [1]; wlam [n] and
 * argmin [1] feed the backward, whose only differentiable input is the roughness */
int robir_spec_prep_fwd(int n, const float* normal, const float* view, const float* rough, const unsigned char* valid,
                        float* ref, float* wl, float* sharp, float* sg_range, float* wlam, int* argmin, void* stream);
int robir_spec_prep_bwd(int n, const float* rough, const float* wlam, const int* argmin, const float* g_sharp,
                        const float* g_sg_range, float* g_rough, void* stream);

/* ---- a10-a12: fused visibility MLP over a pair list: relu(tabA[a]+tabB[b]) -> 3 x (256x256, ReLU) -> sigmoid(z1-z0)
 * replaces the 2M-row VisModel batches of get_diffuse_visibility (model/sg_render.py:157-173) and
 * get_specular_visibility (:259-278).  mask: [rows][4][8] uint32 ReLU sign bits for the backward, or NULL. */
int robir_vis_mlp_fwd(const float* tabA, const float* tabB, const int* rowA, const int* rowB, const int* n_tiles,
                      int max_tiles, const float* Wt1, const float* Wt2, const float* Wt3, const float* b1,
                      const float* b2, const float* b3, const float* wd, const float* bd, float* vis, uint32_t* mask,
                      int sm_count, void* stream);
/* input-gradient of the same MLP w.r.t. the direction (through PE): g_dirs [nB][3] accumulated (zero-initialised). */
int robir_vis_mlp_bwd(const int* rowB, const int* n_tiles, int max_tiles, const float* W1, const float* W2,
                      const float* W3, const float* W0d /*[256][256]: W0[:,63:126] zero padded*/, const float* wd,
                      const float* vis, const float* g_vis, const uint32_t* mask, const float* dirs, float* g_dirs,
                      int sm_count, void* stream);

/* ---- a10-a12, tensor-core engine (tcgen05 + TMEM): same contract as robir_vis_mlp_fwd/bwd over 128-row tiles.
 * terms = 3: fp32-parity mode -- every product is hi*hi + lo*hi + hi*lo on power-of-two-scaled fp16 operands (22
 * mantissa bits per operand, fp32 accumulation); terms = 1: single-pass fp16 fast mode (~1e-4, not the parity mode).
 * Weight images are built once per weight version with robir_tc_pack_layer (same terms): forward = W1, W2, W3
 * (transpose=0); backward = W3^T, W2^T, W1^T (transpose=1) followed by the 64-row image of W0[:, 63:126]^T
 * (n_halves=1).  mask words: [tiles][4][8][128 rows], word w bit (31 - i) = sign of hidden unit 32 w + i. */
int robir_tc_pack_layer(const float* W, int ldw, int N, int K, int transpose, int n_halves, int terms, void* img,
                        void* stream);
int robir_tc_image_bytes(int n_layers256, int n_layers64, int terms);
int robir_vis_tc_fwd(const float* tabA, const float* tabB, const int* rowA, const int* rowB, const int* n_tiles,
                     int max_tiles, const void* img, const float* bias3x256, const float* wd, const float* bd,
                     float* vis, uint32_t* mask, int terms, int sm_count, void* stream);
int robir_vis_tc_bwd(const int* rowB, const int* n_tiles, int max_tiles, const void* img, const float* wd,
                     const float* vis, const float* g_vis, const uint32_t* mask, const float* dirs, float* g_dirs,
                     int terms, int sm_count, void* stream);
/* diagnostics: device buffer [sm_count][8] uint64 that every following vis_tc launch fills with per-CTA stall clocks
 * (issuer: total / wait weights / wait A operand / wait D drained; epilogue: total / wait accumulators / wait X free;
 * producer: wait ring slot); NULL = off */
int robir_tc_debug_buffer(void* buf);
/* unit-test hook: D[128][256] = A[128][256] . W[256][256]^T through the same pipeline (one layer image) */
int robir_tc_selftest(const float* A, const void* img, float* D, int terms, void* stream);

/* ---- 8f rank 4: stage-1 NeuS volume renderer, forward path (neus/volume_render/sdf_render.py) -- the per-ray parts
 * around the network evaluations (robir_sdf_eval + the fused colour chain): up_sample / sample_pdf (:5-82), the sorted
 * merge of cat_z_vals (:85-99), the section midpoints and the alpha compositing of render_core (:141-233) incl. the
 * per-ray depth / accumulation of render_neus (:333-341).  One warp per ray, <= 256 depths per ray.
 * eik_acc [2] (zero-initialised): sum of relax_inside * (|grad| - 1)^2 and of relax_inside. */
int robir_neus_upsample(int B, int n, int n_imp, const float* rays_o, const float* rays_d, const float* z,
                        const float* sdf, float inv_s, float radius, float* new_z, void* stream);
int robir_neus_merge(int B, int n, int m, const float* z, const float* sdf, const float* new_z, const float* new_sdf,
                     float* out_z, float* out_sdf, void* stream);
int robir_neus_midpoints(int B, int n, float sample_dist, const float* rays_o, const float* rays_d, const float* z,
                         float* mid_z, float* pts, void* stream);
int robir_neus_composite(int B, int n, float sample_dist, float inv_s, float cos_anneal, float radius, int white_bkgd,
                         const float* rays_o, const float* rays_d, const float* z, const float* sdf, const float* grad,
                         const float* color, const float* near, const float* far, float* rgb, float* weights,
                         float* acc, float* dist, float* eik_acc, void* stream);

/* ---- a10/a11: weighted per-lobe / per-point means (model/sg_render.py:180-183, :283-294) -------------------------- */
int robir_diffuse_reduce_fwd(int n, int M, int S, const uint32_t* bits, const int* lobe_off, const int* start,
                             const float* vis, const float* w, float* light_vis /*[n][M]*/, void* stream);
int robir_diffuse_reduce_bwd(int n, int M, int S, const uint32_t* bits, const int* lobe_off, const int* start,
                             const float* vis, const float* w, const float* light_vis, const float* g_lv, float* g_vis,
                             float* g_w /*[M*S], zero-initialised*/, void* stream);
int robir_spec_reduce_fwd(int n, int S, int inv, int testing, const int* rowB, const float* vis, const float* w,
                          float* out /*[n]*/, void* stream);
int robir_spec_reduce_bwd(int n, int S, int inv, const int* rowB, const float* vis, const float* w, const float* out,
                          const float* g_out, float* g_vis, float* g_w, void* stream);

/* ---- a9: SG render: render_with_all_sg / render_with_sg (model/sg_render.py:304-565), forward and backward -------- */
typedef struct {
  int n, M, Mi, lin_diff;
  const float *normal, *view, *rough, *albedo, *spec_refl, *lgt, *ind_lgt, *light_vis, *bv_dir, *bv_ind, *ind_integral;
  float *sg_rgb, *sg_spec, *sg_diff, *vis_shadow, *ind_rgb, *ind_spec, *ind_diff, *pre /*[n][9]*/;
  const float *g_sg_rgb, *g_sg_spec, *g_sg_diff, *g_ind_rgb, *g_ind_spec, *g_ind_diff;   /* NULL = zero */
  float *g_lgt /*[M][7] zero-init*/, *g_ind_lgt, *g_light_vis, *g_bv_dir, *g_bv_ind, *g_rough, *g_albedo,
      *g_spec_refl /*[1] zero-init*/, *g_ind_integral;
  float* g_normal;   /* [n][3] or NULL: gradient of the shading normal (CESR renders with normal_net's normals after
                        iteration 1000, training/train_cesr.py:508; the PBR stage passes a detached normal) */
} robir_sg_params;
int robir_sg_render_fwd(const robir_sg_params* p, void* stream);
int robir_sg_render_bwd(const robir_sg_params* p, void* stream);

/* ---- a6/a7/a12/a13: fused small-MLP chains: input encoding + up to 8 Linear(+ReLU|LeakyReLU 0.2) layers in one launch.
 * Replaces the nn.Sequential stacks of SparseAE (model/sg_envmap_material.py:40-99), IndirctIllumNetwork
 * (model/implicit_differentiable_renderer.py:170-222), VisNetwork.forward (:225-258) and the NeuS colour network
 * (model/neus_model.py:489-560).  in_mode: 0 raw rows [n][in_dim]; 1 PE10(x[n][3]); 2 PE10(x) ++ extra[n];
 * 3 IPE10(x, var 1e-5) (model/neus_model.py:25-94); 4 PE10(x[:, :3]) ++ PE10(x[:, 3:]) from x[n][6].
 * Backward: input-gradient chain; per-layer pre-activation gradients are written to L[l].G (weight gradients are
 * G_l^T A_{l-1}, formed by the caller). */
typedef struct {
  const float* Wt;   /* forward  [Kpad][Npad] transposed, zero padded (robir_pack_transpose) */
  const float* Wb;   /* backward [Npad16][Kpad256] row-major, zero padded (robir_pack_pad) */
  const float* bias; /* [Npad] */
  int K, N, Kpad, Npad;
  int act;           /* 0 none, 1 ReLU, 2 LeakyReLU(0.2): applied to this layer's output */
  float* save;       /* post-activation output [n][Npad] (forward: written when non-NULL; backward: read) */
  float* G;          /* backward: pre-activation gradient [n][Npad] or NULL */
} robir_mlp_layer;
typedef struct {
  int n, n_layers, in_mode, in_dim, in_pad;
  const float* x;
  const float* extra;
  const float* noise; /* optional [n][in_dim], added as noise_scale * noise in embedding space (sg_envmap_material.py:83) */
  float noise_scale;
  float* x0_save;     /* embedded input [n][in_pad] or NULL */
  robir_mlp_layer L[8];
  float* out;         /* [n][ldo] */
  int ldo;
  const float* g_out; /* backward: [n][ldo] */
  float* g_x;         /* backward: gradient w.r.t. the embedded input [n][in_pad] or NULL */
  const int* n_active; /* optional device scalar: only rows with (row % seg) < *n_active are evaluated; tiles without
                          such a row write zeros */
  int seg;             /* rows per segment when the batch concatenates equally ordered copies; 0 = n */
} robir_mlp_params;
int robir_pack_pad(const float* W, int N, int K, float* out /*[Np][Kp]*/, int Np, int Kp, void* stream);
int robir_mlp_fwd(const robir_mlp_params* p, int sm_count, void* stream);
int robir_mlp_bwd(const robir_mlp_params* p, int sm_count, void* stream);
/* embedded input of a chain alone -> p->x0_save [n][in_pad] (front end of the tensor-core layer engine) */
int robir_mlp_encode(const robir_mlp_params* p, int sm_count, void* stream);
/* weight / bias gradient of one layer of the chain from its pre-activation gradient G (robir_mlp_bwd) and its input A
 * (x0_save or the previous layer's save): dW [N][K] = G[:, :N]^T A[:, :K], db [N] = column sums (db may be NULL).
 * splits > 1 divides the rows over that many CTAs per 64 x 64 tile (deterministic in-kernel reduction of the partials) */
/* every packed copy of a chain of plain Linear layers (model/sg_envmap_material.py:40-72 nn.Sequential of nn.Linear) in one
 * launch: Wt [Kp][Np] = W^T, Wb [Nb][Kb] = W, bias [Np] = b, all zero padded */
typedef struct {
  const float* W; const float* b;   /* [N][K], [N] */
  float* Wt; float* Wb; float* bias;
  int N, K, Kp, Np, Nb, Kb;
} robir_pack_chain_layer;
typedef struct {
  int n_layers;
  robir_pack_chain_layer L[8];
} robir_pack_chain_params;
int robir_pack_chain(const robir_pack_chain_params* p, void* stream);
int robir_mlp_wgrad(const float* G, int ldg, const float* A, int lda, int n, int N, int K, const int* n_active, int seg,
                    int splits, float* partial /*[splits * tiles * 4160]*/, int* tickets /*[tiles], zero-initialised once*/,
                    float* dW, float* db, void* stream);

/* ---- a6 / a7, tensor-core layer engine (tcgen05 + TMEM, bf16 hi/lo 3-term split = fp32 parity) for the 512-wide
 * encoder / lobe networks: one launch per Linear layer over bf16 hi/lo images (K-major SWIZZLE_128B k-blocks of
 * 128 rows x 64 k, 32 KB each: robir_tl_block_bytes()).  forward: Y = act(A W^T + b); backward: G_prev = (G W) act'(ref).
 * The fp32 rows written by a layer are what robir_mlp_wgrad and the next backward layer consume.
 * Also the engine of the CESR stage's weight-normed chains (section 8f row 1: shadow_net / normal_net,
 * training/train_cesr.py:106-110 = SDFNetwork.forward with multires 0, model/neus_model.py:397-415): act 3 =
 * Softplus(beta = 100), whose derivative is recovered from the saved output (1 - exp(-100 y)); the skip concat is a
 * 512-column row buffer (ld_out > N) that the previous layer fills and robir_tl_pack_rows turns into the next image. */
typedef struct {
  const void* a_img;        /* [ceil(n / 128)][nkb][32 KB] activation (or gradient) image */
  const void* w_img;        /* [ceil(N / 128)][nkb][32 KB] weight image (robir_tl_pack_weight) */
  const float* bias;        /* zero padded to a multiple of 128, or NULL */
  int n, N, nkb, mode, act; /* mode 0 forward / 1 backward; act: this layer's (fwd) or the previous layer's (bwd):
                               0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 Softplus(beta = 100) */
  const float* ref;         /* backward: saved post-activation rows of the previous layer [n][ld_ref], or NULL */
  int ld_ref;
  float* out;               /* fp32 rows [n][ld_out], columns < N, or NULL */
  int ld_out;
  void* out_img;            /* image of the output = next layer's A, [ceil(n / 128)][nkb_out][32 KB], or NULL */
  int nkb_out;
  const int* n_active;      /* as robir_mlp_params */
  int seg;
  int no_fill;              /* 1: leave the outputs of inactive row tiles unwritten instead of zero-filling them (only when
                               every consumer honours the same n_active: the hidden layers of a fixed-capacity chain) */
} robir_tl_params;
int robir_tl_block_bytes(void);
int robir_tl_pack_weight(const float* W, int ldw, int N, int K, int transpose, int col_blocks, int nkb, void* img,
                         void* stream);
int robir_tl_pack_rows(const float* X, int ldx, int n, int K, const float* ref, int ld_ref, int act, int nkb, void* img,
                       void* stream);
int robir_tl_layer(const robir_tl_params* p, void* stream);
/* the same layer as persistent CTAs over 128 x 256 output tiles with two TMEM accumulators (the epilogue of one tile runs
 * under the MMAs of the next) -- identical results, for row counts of several waves (the CESR chains) */
int robir_tl_layer_big(const robir_tl_params* p, int sm_count, void* stream);
/* weight gradients of one layer on the tensor cores, for row counts where robir_mlp_wgrad's fp32 FFMA GEMM dominates
 * (the CESR chains: n_hit x 128 rows).  dW [N][K] = G[:, :N]^T A[:, :K], db [N] = column sums of G (db may be NULL);
 * work holds robir_tl_wgrad_workspace(n, N, K, sm_count) bytes.  G and A are transposed into bf16 hi/lo images whose
 * contraction index (the row) is K-major, a split-K tcgen05 GEMM (3-term hi/lo, fp32 accumulation) fills one 128 x 128
 * tile per CTA and split, and a fixed-order reduction sums the splits: bitwise reproducible.  Replaces, for those
 * shapes, what autograd does for lin.weight / lin.bias of model/neus_model.py:385-417 under training/train_cesr.py:533. */
long long robir_tl_wgrad_workspace(int n, int N, int K, int sm_count);
/* the same weight gradients taken directly from the layer engine's images (no transposing pack): a K-major SWIZZLE_128B
 * block of 128 rows x 64 features is read as an MN-major operand when the contraction runs over the rows.  g_img / a_img:
 * images of G [n][N] and A [n][K] as robir_tl_pack_rows / robir_tl_layer write them (nkb_g / nkb_a k-blocks per row tile);
 * image rows of G at and beyond *n_active (or n) must be zero. */
long long robir_tl_wgrad_mn_workspace(int n, int N, int K, int sm_count);
int robir_tl_wgrad_mn(const void* g_img, int nkb_g, const void* a_img, int nkb_a, int n, int N, int K,
                      const int* n_active, void* work, float* dW, float* db, int sm_count, void* stream);
int robir_tl_wgrad(const float* G, int ldg, const float* A, int lda, int n, int N, int K,
                   const int* n_active /* device row count of a fixed-capacity batch (rows beyond it are zero), or NULL */,
                   void* work, float* dW, float* db, int sm_count, void* stream);
/* ---- a1: hit compaction of a fixed-capacity ray batch (device-side counterpart of the boolean indexing at
 * implicit_differentiable_renderer.py:341-347): hits first (stable), misses after; pos / order int64 [N], n_act [1],
 * valid [N] (slot < n_act), pts [N][3] = hit points in slot order (0 for misses), view [N][3] = -dirs in slot order */
int robir_compact_hits(int N, const unsigned char* hit, const float* points, const float* dirs, long long* pos,
                       long long* order, int* n_act, unsigned char* valid, float* pts, float* view, void* stream);

/* ---- a7: SparseAE glue of the BRDF auto-encoder (model/sg_envmap_material.py:74-94, 214-232): the decoder's doubled
 * latent batch [sigmoid(z); sigmoid(z) + 0.01 noise] and the output head (sigmoid + affine -> albedo / roughness /
 * metallic and their random_xi twins), one launch each way.  Null upstream gradients count as zero. ---------------- */
int robir_latent_pair_fwd(int n, const float* z /*[n][32]*/, const float* noise, float* out /*[2n][32]*/, void* stream);
int robir_latent_pair_bwd(int n, const float* z, const float* g /*[2n][32]*/, float* g_z, void* stream);
int robir_brdf_head_fwd(int n, const float* y2 /*[2n][5]*/, float* albedo, float* rough, float* metal, float* xi_albedo,
                        float* xi_rough, float* xi_metal, void* stream);
int robir_brdf_head_bwd(int n, const float* y2, const float* g_albedo, const float* g_rough, const float* g_metal,
                        const float* g_xi_albedo, const float* g_xi_rough, const float* g_xi_metal, float* g_y2,
                        void* stream);

/* ---- a6: lobe decoding of IndirctIllumNetwork (implicit_differentiable_renderer.py:207-219): raw [total][6] ->
 * [axis(theta = 2 pi sigmoid, phi = pi sigmoid), 30 sigmoid + 0.1, relu x3] [total][7], total = points x lobes --------- */
int robir_decode_lobes_fwd(int total, const float* raw, float* sgs, void* stream);
int robir_decode_lobes_bwd(int total, const float* raw, const float* g_sgs, float* g_raw, void* stream);

/* ---- loss epilogue: model/loss.py:61-125 (InvLoss: masked L1/L2 on the ACES tone-mapped radiance,
 * latent-smooth L1, KL sparsity of the BRDF latent), model/color_correction.py:31-59 (hdr2ldr with the learnable
 * exposure shift), training/train_pbr.py:313-346 (white_loss; loss = rgb + kl + 0.1 smooth + white).  One launch
 * computes the value and every input gradient. ------------------------------------------------------------------- */
typedef struct {
  int N, n_lat, M, l2;
  const float* sg_rgb; const float* indir_rgb;   /* [N][3], row strides ld_sg / ld_ind (floats) */
  int ld_sg, ld_ind;
  const float* gt;                                /* [N][3] */
  const unsigned char* mask;                      /* [N] network_object_mask & object_mask (ray order) */
  const unsigned char* hit;                       /* optional [N] network_object_mask (ray order), used with order */
  const long long* order;                         /* optional [N]: input row i belongs to ray order[i] (hit rays compacted
                                                     to the front); gt / mask / hit stay in ray order */
  const float* adapt_illum;                       /* [1] */
  const float* albedo; const float* albedo_r;     /* [N][3], strides ld_alb / ld_albr */
  int ld_alb, ld_albr;
  const float* rough; const float* rough_r;       /* [N] (column 0), strides ld_r / ld_rr */
  int ld_r, ld_rr;
  const float* z;                                 /* [n_lat][32] */
  const unsigned char* z_valid;                   /* [n_lat] or NULL */
  const float* lgt;                               /* [M][7] */
  float w_rgb, w_kl, w_smooth, rho;
  float* losses;                                  /* [5] total, sg_rgb_loss, kl, smooth, white */
  float* g_pred; float* g_adapt; float* g_albedo; float* g_albedo_r; float* g_rough; float* g_rough_r;
  float* g_z; float* g_lgt;
} robir_loss_params;
int robir_pbr_loss(const robir_loss_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROBIR_B200_H_ */
