"""Random draws of the hot path.  The reference draws every hot-path random tensor on the CPU generator and copies it
to the GPU (model/sg_render.py:135-136,224-225; model/sg_envmap_material.py:81,83).  Modes:

  "cpu"     (default) identical call order / shapes on torch's CPU generator -> same numbers as the reference under the
            same ``torch.manual_seed``; costs a host draw + H2D copy per call, like the reference;
  "device"  draw on the CUDA generator (no host round trip; different numbers);
  "replay"  pop tensors from a tape (parity tests feed the reference's recorded randoms).
"""
import contextlib

import torch

_mode = "cpu"
_tape = None
_record = None
_record_on_device = False


def set_mode(mode):
    global _mode
    assert mode in ("cpu", "device", "replay")
    _mode = mode


@contextlib.contextmanager
def replay(tensors):
    """Feed the given list of tensors, in order, to the next draws."""
    global _mode, _tape
    old = (_mode, _tape)
    _mode, _tape = "replay", list(tensors)
    try:
        yield
    finally:
        _mode, _tape = old


@contextlib.contextmanager
def record(on_device=False):
    """Collect every drawn tensor, in call order.  on_device=False: host copies (one D2H per draw; not capturable).
    on_device=True: the drawn device tensors themselves, nothing else changes -- usable around a CUDA-graph capture:
    the tensors then belong to the graph and hold, after each replay, the numbers that replay used (the parity tests
    read the randoms of the benchmarked graph step this way)."""
    global _record, _record_on_device
    _record, _record_on_device = [], on_device
    try:
        yield _record
    finally:
        _record, _record_on_device = None, False


def _draw(fn, shape, device):
    shape = tuple(int(s) for s in shape)
    if _mode == "replay":
        t = _tape.pop(0)
        assert tuple(t.shape) == shape, "replayed random tensor has shape %s, expected %s" % (tuple(t.shape), shape)
        out = t.to(device=device, dtype=torch.float32)
    elif _mode == "device":
        out = fn(shape, device=device)
    else:
        out = fn(shape).to(device, non_blocking=True)
    if _record is not None:
        _record.append(out if _record_on_device else out.detach().cpu().clone())
    return out


def rand(shape, device):
    return _draw(torch.rand, shape, device)


def rand_pairs(n, S, device):
    """theta, phi for two consecutive get_specular_visibility calls, stacked: ([2n, S], [2n, S]).  Host / replay modes
    draw theta_1, phi_1, theta_2, phi_2 in the reference's order; the device mode draws each stack at once."""
    if _mode == "device" and (_record is None or _record_on_device):
        return rand((2 * n, S), device), rand((2 * n, S), device)
    u = [rand((n, S), device) for _ in range(4)]
    return torch.cat([u[0], u[2]], 0), torch.cat([u[1], u[3]], 0)


def randn(shape, device):
    return _draw(torch.randn, shape, device)


def uniform(shape, device):
    """torch.empty(shape).uniform_(0, 1) of the reference (model/ray_tracing.py:305): same generator stream as rand."""
    return _draw(lambda s, **kw: torch.empty(s, **kw).uniform_(0.0, 1.0), shape, device)
