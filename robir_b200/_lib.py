"""ctypes binding of the C-ABI shared library ``librobir_b200.so`` (declared in include/robir_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or was not built for this GPU every entry point raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librobir_b200.so")
_lib = None
_sm_count = {}


class RobirError(RuntimeError):
    pass


class SgParams(Structure):
    _fields_ = [("n", c_int), ("M", c_int), ("Mi", c_int), ("lin_diff", c_int)] + [
        (k, c_void_p) for k in (
            "normal", "view", "rough", "albedo", "spec_refl", "lgt", "ind_lgt", "light_vis", "bv_dir", "bv_ind",
            "ind_integral", "sg_rgb", "sg_spec", "sg_diff", "vis_shadow", "ind_rgb", "ind_spec", "ind_diff", "pre",
            "g_sg_rgb", "g_sg_spec", "g_sg_diff", "g_ind_rgb", "g_ind_spec", "g_ind_diff", "g_lgt", "g_ind_lgt",
            "g_light_vis", "g_bv_dir", "g_bv_ind", "g_rough", "g_albedo", "g_spec_refl", "g_ind_integral", "g_normal")]


class SdfParams(Structure):
    _fields_ = [("pts", c_void_p), ("n", c_int), ("in_scale", c_float), ("sdf_scale", c_float),
                ("feat_scale", c_float), ("Wt", c_void_p * 8), ("bias", c_void_p * 8), ("w8_sdf", c_void_p),
                ("b8", c_void_p), ("Wt8_feat", c_void_p), ("sdf", c_void_p), ("grad", c_void_p), ("feat", c_void_p),
                ("n_active", c_void_p)]


class SdfTcParams(Structure):
    _fields_ = [("pts", c_void_p), ("n", c_int), ("in_scale", c_float), ("sdf_scale", c_float),
                ("feat_scale", c_float), ("img", c_void_p), ("bias", c_void_p), ("w8_sdf", c_void_p), ("b8", c_void_p),
                ("sdf", c_void_p), ("grad", c_void_p), ("feat", c_void_p), ("n_active", c_void_p)]


class SdfNet(Structure):
    _fields_ = [("Wt", c_void_p * 8), ("bias", c_void_p * 8), ("w8_sdf", c_void_p), ("b8", c_void_p)]


class SphereTraceParams(Structure):
    _fields_ = [("net", SdfNet), ("N", c_int), ("o_div", c_int), ("cam_loc", c_void_p), ("ray_dirs", c_void_p),
                ("object_mask", c_void_p), ("in_scale", c_float), ("out_scale", c_float), ("radius", c_float),
                ("sdf_threshold", c_float), ("line_search_step", c_float), ("line_step_iters", c_int),
                ("sphere_tracing_iters", c_int), ("n_steps", c_int), ("n_secant_steps", c_int), ("training", c_int)] + [
        (k, c_void_p) for k in ("uniform_steps", "points", "net_mask", "dists", "acc_s", "acc_e", "min_dis", "max_dis",
                                "flags", "samp_list", "sec_list", "sec_state", "min_list", "vals", "counters")]


class TlParams(Structure):
    _fields_ = [("a_img", c_void_p), ("w_img", c_void_p), ("bias", c_void_p), ("n", c_int), ("N", c_int), ("nkb", c_int),
                ("mode", c_int), ("act", c_int), ("ref", c_void_p), ("ld_ref", c_int), ("out", c_void_p),
                ("ld_out", c_int), ("out_img", c_void_p), ("nkb_out", c_int), ("n_active", c_void_p), ("seg", c_int),
                ("no_fill", c_int)]


class LossParams(Structure):
    _fields_ = [("N", c_int), ("n_lat", c_int), ("M", c_int), ("l2", c_int), ("sg_rgb", c_void_p),
                ("indir_rgb", c_void_p), ("ld_sg", c_int), ("ld_ind", c_int), ("gt", c_void_p), ("mask", c_void_p),
                ("hit", c_void_p), ("order", c_void_p), ("adapt_illum", c_void_p), ("albedo", c_void_p), ("albedo_r", c_void_p), ("ld_alb", c_int),
                ("ld_albr", c_int), ("rough", c_void_p), ("rough_r", c_void_p), ("ld_r", c_int), ("ld_rr", c_int),
                ("z", c_void_p), ("z_valid", c_void_p), ("lgt", c_void_p), ("w_rgb", c_float), ("w_kl", c_float),
                ("w_smooth", c_float), ("rho", c_float)] + [
        (k, c_void_p) for k in ("losses", "g_pred", "g_adapt", "g_albedo", "g_albedo_r", "g_rough", "g_rough_r", "g_z",
                                "g_lgt")]


class MlpLayer(Structure):
    _fields_ = [("Wt", c_void_p), ("Wb", c_void_p), ("bias", c_void_p), ("K", c_int), ("N", c_int), ("Kpad", c_int),
                ("Npad", c_int), ("act", c_int), ("save", c_void_p), ("G", c_void_p)]


class PackChainLayer(Structure):
    _fields_ = [("W", c_void_p), ("b", c_void_p), ("Wt", c_void_p), ("Wb", c_void_p), ("bias", c_void_p),
                ("N", c_int), ("K", c_int), ("Kp", c_int), ("Np", c_int), ("Nb", c_int), ("Kb", c_int)]


class PackChainParams(Structure):
    _fields_ = [("n_layers", c_int), ("L", PackChainLayer * 8)]


class MlpParams(Structure):
    _fields_ = [("n", c_int), ("n_layers", c_int), ("in_mode", c_int), ("in_dim", c_int), ("in_pad", c_int),
                ("x", c_void_p), ("extra", c_void_p), ("noise", c_void_p), ("noise_scale", c_float),
                ("x0_save", c_void_p), ("L", MlpLayer * 8), ("out", c_void_p), ("ldo", c_int), ("g_out", c_void_p),
                ("g_x", c_void_p), ("n_active", c_void_p), ("seg", c_int)]


class OctreeView(Structure):
    _fields_ = [("nodes", c_void_p), ("grid", c_void_p), ("gx", c_int), ("gy", c_int), ("gz", c_int),
                ("n_nodes", c_int), ("rminx", c_float), ("rminy", c_float), ("rminz", c_float), ("rsizex", c_float),
                ("rsizey", c_float), ("rsizez", c_float)]


class OctCastParams(Structure):
    _fields_ = [("view", OctreeView), ("sdf_grad", c_void_p), ("rays_o", c_void_p), ("rays_d", c_void_p),
                ("K", c_int), ("o_div", c_int), ("max_iter", c_int), ("eps", c_float), ("refine_limit", c_float),
                ("last_node_sdf", c_float), ("state_t", c_void_p), ("state_ptr", c_void_p), ("out_t", c_void_p),
                ("out_x", c_void_p), ("out_hit", c_void_p), ("counters", c_void_p)]


# name -> argtypes (all functions return int status except where noted)
_P, _I, _F = c_void_p, c_int, c_float
_SIGNATURES = {
    "robir_pack_transpose": [_P, _I, _I, _I, _I, _P, _I, _I, _F, _P],
    "robir_pack_window": [_P, _I, _I, _I, _I, _P, _I, _I, _P],
    "robir_pack_wn_transpose": [_P, _P, _I, _I, _I, _I, _P, _I, _I, _P],
    "robir_pack_wn_row": [_P, _P, _I, _I, _P, _P],
    "robir_pe_linear": [_P, _I, _P, _P, _P, _P],
    "robir_sample_dirs_fwd": [_I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P],
    "robir_sample_dirs_bwd": [_I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_diffuse_rows": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_spec_rows": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "robir_spec_prep_fwd": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_spec_prep_bwd": [_I, _P, _P, _P, _P, _P, _P, _P],
    "robir_tc_pack_layer": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "robir_tc_image_bytes": [_I, _I, _I],
    "robir_vis_tc_fwd": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "robir_vis_tc_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "robir_tc_selftest": [_P, _P, _P, _I, _P],
    "robir_tc_debug_buffer": [_P],
    "robir_neus_upsample": [_I, _I, _I, _P, _P, _P, _P, _F, _F, _P, _P],
    "robir_neus_merge": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "robir_neus_midpoints": [_I, _I, _F, _P, _P, _P, _P, _P, _P],
    "robir_neus_composite": [_I, _I, _F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_pack_pad": [_P, _I, _I, _P, _I, _I, _P],
    "robir_pack_chain": [POINTER(PackChainParams), _P],
    "robir_mlp_fwd": [POINTER(MlpParams), _I, _P],
    "robir_mlp_bwd": [POINTER(MlpParams), _I, _P],
    "robir_vis_mlp_fwd": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "robir_vis_mlp_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "robir_diffuse_reduce_fwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "robir_diffuse_reduce_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_spec_reduce_fwd": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "robir_spec_reduce_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_sg_render_fwd": [POINTER(SgParams), _P],
    "robir_sg_render_bwd": [POINTER(SgParams), _P],
    "robir_sdf_eval": [POINTER(SdfParams), _I, _P],
    "robir_sdf_tc": [POINTER(SdfTcParams), _I, _P],
    "robir_camera_rays": [_I, _P, _P, _P, _P, _P],
    "robir_octree_cast": [POINTER(OctCastParams), _I, _P],
    "robir_octree_counters_len": [],
    "robir_sphere_trace": [POINTER(SphereTraceParams), _I, _P],
    "robir_sphere_trace_launches": [_I],
    "robir_mlp_wgrad": [_P, _I, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _P, _P, _P],
    "robir_compact_hits": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_latent_pair_fwd": [_I, _P, _P, _P, _P],
    "robir_latent_pair_bwd": [_I, _P, _P, _P, _P],
    "robir_brdf_head_fwd": [_I, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_brdf_head_bwd": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "robir_decode_lobes_fwd": [_I, _P, _P, _P],
    "robir_decode_lobes_bwd": [_I, _P, _P, _P, _P],
    "robir_tl_block_bytes": [],
    "robir_tl_pack_weight": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "robir_tl_pack_rows": [_P, _I, _I, _I, _P, _I, _I, _I, _P, _P],
    "robir_tl_layer": [POINTER(TlParams), _P],
    "robir_tl_layer_big": [POINTER(TlParams), _I, _P],
    "robir_tl_wgrad_workspace": [_I, _I, _I, _I],
    "robir_tl_wgrad_mn_workspace": [_I, _I, _I, _I],
    "robir_tl_wgrad_mn": [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P],
    "robir_tl_wgrad": [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P],
    "robir_mlp_encode": [POINTER(MlpParams), _I, _P],
    "robir_pbr_loss": [POINTER(LossParams), _P],
    "robir_device_info": [POINTER(c_int), POINTER(c_int), POINTER(c_int)],
    "robir_abi_version": [],
}
EXPORTED = sorted(list(_SIGNATURES) + ["robir_last_error"])


# kernels launched per C call (for bench.py's gpu_launches claim); everything not listed launches exactly one
_KERNELS_PER_CALL = {"robir_diffuse_rows": 3, "robir_sphere_trace": 7, "robir_sphere_trace_launches": 0, "robir_octree_counters_len": 0, "robir_device_info": 0,
                     "robir_tc_image_bytes": 0, "robir_tl_block_bytes": 0,
                     "robir_tl_wgrad_workspace": 0, "robir_tl_wgrad": 4,
                     "robir_tl_wgrad_mn_workspace": 0, "robir_tl_wgrad_mn": 2,
                     "robir_abi_version": 0, "robir_last_error": 0}
launch_count = 0


class _Counting:
    """Thin proxy over the CDLL handle that counts kernel launches issued through the C ABI."""

    def __init__(self, handle):
        self._h = handle

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        k = _KERNELS_PER_CALL.get(name, 1)
        if k == 0:
            return fn

        def call(*a):
            global launch_count
            launch_count += k
            return fn(*a)
        return call


def lib():
    """Load (once) and return the ctypes handle; raises RobirError when the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RobirError("robir_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                             "there is no CPU fallback" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_longlong if name.endswith("_workspace") else c_int
        handle.robir_last_error.restype = c_char_p
        handle.robir_last_error.argtypes = []
        _lib = _Counting(handle)
    return _lib


def check(status):
    if status != 0:
        raise RobirError("robir_b200: %s" % lib().robir_last_error().decode())


def require_cuda(t):
    if not t.is_cuda:
        raise RobirError("robir_b200 ops need CUDA tensors (got %s); there is no CPU fallback" % t.device)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    require_cuda(t)
    if not t.is_contiguous():
        raise RobirError("robir_b200: non-contiguous tensor passed to the C ABI")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def sm_count():
    dev = torch.cuda.current_device()
    if dev not in _sm_count:
        a, b, c = c_int(), c_int(), c_int()
        check(lib().robir_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        if b.value < 10:
            raise RobirError("robir_b200 is built for sm_100a only; device reports sm_%d%d" % (b.value, c.value))
        _sm_count[dev] = a.value
    return _sm_count[dev]


def f32(t):
    return t.detach().contiguous().float()
