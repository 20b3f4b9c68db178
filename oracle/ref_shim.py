"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Makes the *unmodified* reference tree (/root/reference, read-only) importable and runnable on CPU in the
build container so that (a) the oracle restatement in ``oracle/robir_oracle.py`` can be validated against
the real thing and (b) golden vectors under ``tests/golden/`` can be generated (``tests/golden/make_golden.py``).

/root/reference does not exist on the GPU box; there the byte-for-byte staged copy ``oracle/_ref/`` (made by
``oracle/stage_ref.py`` from ``__graft_entry__.build()``, git-ignored, shipped by gpurun) is used instead, by
``bench.py --impl reference[-cuda]`` and by the ``-m gpu`` drop-in test of ``robir_b200.install()``.
Recipe: SURVEY.md Appendix B.

What it does (nothing under the reference tree is edited):
  * stub modules for the reference's missing third-party imports (gin, imageio, torch_scatter, pyhocon, ...);
    ``torch_scatter.scatter_min`` (utils/octree.py:591) is restated as a segment-amin;
  * device="cpu" (default): ``.cuda()`` becomes a no-op and ``device='cuda'`` kwargs are rewritten to CPU;
    device="cuda": nothing is rewritten -- the reference runs eagerly on the GPU exactly as its users run it;
  * registers <reference>/datasets under the name ``datasets`` (HF ``datasets`` shadows it).
"""
import os
import sys
import tempfile
import types

import torch

def _find_root():
    cands = [os.environ.get("ROBIR_REFERENCE"), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "model")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()
_installed = False
DEVICE = "cpu"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _passthrough_decorator(*a, **k):
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return lambda f: f


def _scatter_min(src, index):
    # torch_scatter.scatter_min semantic used at utils/octree.py:591: per-segment minimum of src.
    if index.numel() == 0:
        return src.new_zeros((0,)), None
    out = torch.full((int(index.max()) + 1,), torch.iinfo(src.dtype).max, dtype=src.dtype, device=src.device)
    return out.scatter_reduce(0, index, src, "amin", include_self=True), None


def install(device="cpu"):
    """Idempotently install the shim and put the reference on sys.path.  device: "cpu" (rewrite every CUDA placement
    to the host) or "cuda" (leave the reference's own .cuda() calls alone); fixed by the first call in a process."""
    global _installed, DEVICE
    if _installed:
        if device != DEVICE and device == "cuda":
            raise RuntimeError("ref_shim was already installed in CPU mode in this process")
        return
    DEVICE = device
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    gin = _mod("gin", configurable=_passthrough_decorator, register=_passthrough_decorator,
               add_config_file_search_path=lambda *a, **k: None)
    gin.config = _mod("gin.config", external_configurable=lambda *a, **k: None)
    io = _mod("imageio", imread=None, imwrite=None)
    io.plugins = _mod("imageio.plugins")
    io.plugins.freeimage = _mod("imageio.plugins.freeimage", download=lambda: None)
    _mod("torch_scatter", scatter_min=_scatter_min)
    for n in ["matplotlib", "matplotlib.pyplot", "trimesh", "xatlas", "glfw", "OpenGL"]:
        _mod(n)
    _mod("OpenGL.GL", GL_TRIANGLES=4)
    _mod("OpenGL.GL.shaders", compileShader=None, compileProgram=None)
    _mod("pyhocon", ConfigFactory=None)
    _mod("tensorboardX", SummaryWriter=None)

    def _nan_guard():
        raise RuntimeError("reference NaN guard (ipdb.set_trace) hit")

    _mod("ipdb", set_trace=_nan_guard)
    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REF_ROOT, "datasets")]
    sys.modules["datasets"] = ds

    if device == "cpu":
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        from torch import storage as _storage
        _storage._StorageBase.cuda = lambda self, *a, **k: self
        torch.UntypedStorage.cuda = lambda self, *a, **k: self

        from torch.overrides import TorchFunctionMode

        class _CpuDevice(TorchFunctionMode):
            def __torch_function__(self, func, types_, args=(), kwargs=None):
                kwargs = kwargs or {}
                dev = kwargs.get("device")
                if dev is not None and "cuda" in str(dev):
                    kwargs["device"] = "cpu"
                return func(*args, **kwargs)

        _CpuDevice().__enter__()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import utils.octree as uo

    if device == "cpu":
        def _octree_cuda(self):
            self.device = "cpu"
            return self

        uo.Octree.cuda = _octree_cuda
    _installed = True


class DictConf(dict):
    """Stand-in for the pyhocon ConfigTree the reference constructors expect."""

    def _get(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def get_bool(self, k):
        return bool(self._get(k))

    def get_int(self, k):
        return int(self._get(k))

    def get_float(self, k):
        return float(self._get(k))

    def get_config(self, k):
        return DictConf(self._get(k))


def hotdog_model_conf(num_lgt_sgs=128, use_octree=True, n_steps=100):
    """The model{} block of confs_sg/hotdog.conf:65-123 as a dict (values restated, not parsed)."""
    return DictConf(
        gamma=1.0, hdr_mode=0, use_neus=True, use_octree=use_octree, feature_vector_size=256,
        implicit_network=dict(d_in=3, d_out=1, dims=[512] * 8, geometric_init=True, bias=0.6, skip_in=[4],
                              weight_norm=True, multires=6),
        rendering_network=dict(mode="idr", d_in=9, d_out=3, dims=[512] * 4, weight_norm=True, multires_view=4),
        indirect_illum_network=dict(multires=10, dims=[512] * 4, num_lgt_sgs=24),
        visibility_network=dict(points_multires=10, dirs_multires=10, dims=[256] * 4),
        envmap_material_network=dict(multires=10, brdf_encoder_dims=[512] * 4, brdf_decoder_dims=[128, 128],
                                     num_lgt_sgs=num_lgt_sgs, upper_hemi=False, specular_albedo=0.05, latent_dim=32),
        ray_tracer=dict(object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=3,
                        sphere_tracing_iters=10, n_steps=n_steps, n_rootfind_steps=32),
    )


def build_reference_model(neus_state_dict, num_lgt_sgs=128, use_octree=True, n_steps=100, seed=0, device="cpu"):
    """Construct the reference IDRNetwork from a stage-1 NeuS state dict (our synthetic checkpoint); on the host
    (device="cpu") or -- the caller then does ``model.cuda()`` like the runners, train_pbr.py:92-93 -- for the GPU."""
    install(device)
    import confs_sg.env_path as env_path
    tmp = tempfile.mkdtemp(prefix="robir_neus_")
    torch.save({"global_step": 0, "model": neus_state_dict}, os.path.join(tmp, "000000.tar"))
    env_path.set_path(tmp, 0)
    from model.implicit_differentiable_renderer import IDRNetwork
    torch.manual_seed(seed)
    return IDRNetwork(hotdog_model_conf(num_lgt_sgs, use_octree, n_steps))


def reference_neus_state_dict(seed=0):
    install()
    from model.neus_model import NeuSModel
    torch.manual_seed(seed)
    return NeuSModel(mode="idr", hashing=False, embed="PE").state_dict()


def bind_pbr_runner(model, no_normal=True, is_training=True):
    """A bare PBRTrainRunner carrying only what get_sg_render (training/train_pbr.py:348-396) reads."""
    install()
    from training.train_pbr import PBRTrainRunner
    runner = PBRTrainRunner.__new__(PBRTrainRunner)
    runner.model = model
    runner.train_spec = True
    runner.is_training = is_training
    runner.no_normal = no_normal
    model.get_sg_render = runner.get_sg_render
    return runner


def bind_cesr_runner(model, cur_iter=600, white_light=False, explore_iter=1000, proj_iter=0, explore_smooth=0.1,
                     explore_kl=1.0, proj_smooth=0.01, proj_kl=0.01, is_training=True, seed=0):
    """A bare ClusteredAlbedoTrainRunner carrying what get_sg_render / pbr_step (training/train_cesr.py:387-430,465-559)
    read; shadow_net / normal_net are built exactly as at :102-110 (their own default initialisation, torch seed)."""
    install()
    from training.train_cesr import ClusteredAlbedoTrainRunner
    from model.neus_model import SDFNetwork
    from model.embedder import get_embedder
    runner = ClusteredAlbedoTrainRunner.__new__(ClusteredAlbedoTrainRunner)
    runner.model = model
    runner.train_spec = True
    runner.is_training = is_training
    runner.cur_iter = cur_iter
    runner.white_light = white_light
    runner.conf = DictConf(train=dict(argmax_vis=False, explore_iter=explore_iter, proj_iter=proj_iter,
                                      explore_smooth=explore_smooth, explore_kl=explore_kl, proj_smooth=proj_smooth,
                                      proj_kl=proj_kl))
    torch.manual_seed(seed)
    runner.shadow_embed, in_dim = get_embedder(10)
    runner.shadow_net = SDFNetwork(in_dim + 128, 2, 512, 8, [4], 0)
    runner.normal_net = SDFNetwork(in_dim, 3, 512, 8, [4], 0)
    model.get_sg_render = runner.get_sg_render
    return runner


def load_stage1_renderer():
    """The UNMODIFIED neus/volume_render/sdf_render.py as a module.  Its two star imports (``misc.utils`` / ``misc.defs``,
    which drag in absl, PIL and the stage-1 package layout) are satisfied by stub modules that export exactly the names
    the file uses: torch, F, np, gin (the pass-through stub of install()) and the Rays / IComp / ISDF definitions of
    neus/misc/defs.py:8-37 (a namedtuple and two interface classes, restated)."""
    install()
    import collections
    import importlib.util
    import numpy as np
    import torch.nn.functional as F
    Rays = collections.namedtuple('Rays', ('origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'near', 'far'))
    IComp = type("IComp", (), {})
    ISDF = type("ISDF", (IComp,), {})
    saved = {k: sys.modules.get(k) for k in ("misc", "misc.utils", "misc.defs")}
    pkg = _mod("misc")
    pkg.__path__ = []
    names = dict(torch=torch, F=F, np=np, gin=sys.modules["gin"])
    _mod("misc.utils", __all__=list(names), **names)
    defs = dict(Rays=Rays, IComp=IComp, ISDF=ISDF)
    _mod("misc.defs", __all__=list(defs), **defs)
    try:
        spec = importlib.util.spec_from_file_location(
            "robir_ref_stage1_sdf_render", os.path.join(REF_ROOT, "neus", "volume_render", "sdf_render.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class ReplayRandom:
    """Context manager: record (mode='record') or replay (mode='replay') torch.rand / torch.randn /
    Tensor.uniform_ draws, in call order, so reference and oracle/product see identical randoms."""

    def __init__(self, tape=None):
        self.tape = [] if tape is None else list(tape)
        self.replay = tape is not None
        self._pos = 0

    def _wrap(self, name, orig):
        def f(*a, **k):
            if self.replay:
                out = self.tape[self._pos][1].clone()
                self._pos += 1
                return out
            out = orig(*a, **k)
            self.tape.append((name, out.clone()))
            return out
        return f

    def __enter__(self):
        self._orig = (torch.rand, torch.randn)
        torch.rand = self._wrap("rand", self._orig[0])
        torch.randn = self._wrap("randn", self._orig[1])
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._orig
        return False
