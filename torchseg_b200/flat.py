"""Flat fp32 parameter / gradient buffers shared by the fused SGD and the DDP gradient all-reduce.

Ownership: a parameter's gradient can live in exactly ONE flat buffer. The layout created by the first owner (the fused
optimiser or the DDP shim) is registered per parameter-id tuple; a DDP shim built over the SAME parameter set re-uses it;
a new fused optimiser takes the parameters over (the old owner's next zero_grad / step then raises); any other overlap
is refused loudly — silently re-pointing `.grad` would leave the first owner reading a stale buffer of zeros."""
import weakref

import torch

_layouts = {}     # tuple(id(p)) -> (flat, spans, params)
_owner_of = {}    # id(p) -> layout key


def flatten(params, attr, like_data=True):
    """Make every tensor `getattr(p, attr)` (attr in {'data','grad'}) a view into ONE contiguous fp32 buffer,
    preserving each parameter's strides (KRSC conv weights stay KRSC). Returns (flat, spans)."""
    dev = params[0].device
    # every span starts on a 32-byte boundary: float4 / red.v4 paths need 16 B, and the bf16 mirror of the parameter
    # buffer (optim.FlatPack) must keep every weight 16-byte aligned for the TMA descriptors; the padding stays zero
    n = sum((p.numel() + 7) // 8 * 8 for p in params)
    flat = torch.zeros(n, dtype=torch.float32, device=dev)
    spans = []
    off = 0
    for p in params:
        k = (p.numel() + 7) // 8 * 8
        view = torch.as_strided(flat, p.shape, p.stride(), off)
        if attr == "data":
            view.copy_(p.data)
            p.data = view
        else:
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            p._tsb_direct = True   # kernels may accumulate straight into this persistent view (ops._direct_grad)
        spans.append((off, off + k))
        off += k
    return flat, spans


def _repoint(params, flat, spans):
    """re-establish the views if someone set a grad to None / replaced it (zero_grad(set_to_none=True), a hook)"""
    base = flat.data_ptr()
    for p, (lo, hi) in zip(params, spans):
        if p.grad is None or p.grad.data_ptr() != base + 4 * lo:
            view = torch.as_strided(flat, p.shape, p.stride(), lo)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            p._tsb_direct = True


def _alive(key):
    """the layout registered under `key` if every parameter it was built over is still alive, else None (and purge it)"""
    lay = _layouts.get(key)
    if lay is None:
        return None
    ps = [r() for r in lay[2]]
    if any(q is None for q in ps) or tuple(id(q) for q in ps) != key:
        _layouts.pop(key, None)
        for i in key:
            if _owner_of.get(i) == key:
                _owner_of.pop(i, None)
        return None
    return lay[0], lay[1], ps


def layout_of(params):
    """the registered layout (flat, spans, params-in-layout-order) covering exactly this parameter SET, or None"""
    ids = set(id(p) for p in params)
    keys = set(_owner_of.get(i) for i in ids)
    if len(keys) != 1 or None in keys:
        return None
    key = next(iter(keys))
    lay = _alive(key)
    if lay is None or set(key) != ids:
        return None
    by_id = dict((id(p), p) for p in params)
    if any(by_id[id(q)] is not q for q in lay[2]):
        return None
    return lay


def ensure_flat_grads(params, take_over=False):
    """Idempotent: returns (flat_grad, spans) for this exact parameter LIST. Owners must use the returned buffer (and
    refresh their cached reference from it) rather than assume an earlier one is still current.
    take_over=True (owner construction only): an older layout over some of these parameters is dropped — its owner then
    fails loudly (this function raises for it) instead of silently reading a buffer nobody writes."""
    key = tuple(id(p) for p in params)
    lay = _alive(key)
    if lay is not None and all(a is b for a, b in zip(lay[2], params)):
        _repoint(params, lay[0], lay[1])
        return lay[0], lay[1]
    clash = 0
    for p in params:
        other = _owner_of.get(id(p))
        if other is not None:
            olay = _alive(other)
            if olay is not None and any(q is p for q in olay[2]):
                clash += 1
    if clash and take_over:
        release(params)
    elif clash:
        raise RuntimeError(
            "flat gradient layout conflict: %d parameter(s) already belong to another flat gradient buffer (a fused "
            "optimiser / DistributedDataParallel built over a different parameter list). Build both over the same "
            "parameters (DistributedDataParallel re-uses the optimiser's layout when the sets are equal), or call "
            "torchseg_b200.flat.release(params) first." % clash)
    flat, spans = flatten(params, "grad")
    _layouts[key] = (flat, spans, [weakref.ref(p) for p in params])
    for p in params:
        _owner_of[id(p)] = key
    return flat, spans


def release(params):
    """forget the layouts that contain any of these parameters (their .grad tensors stay valid views)"""
    for p in params:
        key = _owner_of.pop(id(p), None)
        if key is not None:
            _layouts.pop(key, None)


def check_inside(params, flat, spans):
    """cheap guard used right before a kernel consumes `flat`: first and last parameter's .grad must be views of it"""
    base = flat.data_ptr()
    for idx in (0, len(params) - 1):
        p, lo = params[idx], spans[idx][0]
        if p.grad is None or p.grad.data_ptr() != base + 4 * lo:
            raise RuntimeError("parameter %d's .grad is not a view of this owner's flat gradient buffer" % idx)
