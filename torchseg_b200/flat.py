"""Flat fp32 parameter / gradient buffers shared by the fused SGD and the DDP gradient all-reduce."""
import torch


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


def ensure_flat_grads(params):
    """idempotent: returns (flat_grad, spans) for this exact parameter list"""
    key = tuple(id(p) for p in params)
    reg = getattr(ensure_flat_grads, "_reg", None)
    if reg is not None and reg[0] == key:
        flat, spans = reg[1], reg[2]
        for p, (lo, hi) in zip(params, spans):  # re-point if someone set grads to None
            if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + 4 * lo:
                view = torch.as_strided(flat, p.shape, p.stride(), lo)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
                p._tsb_direct = True
        return flat, spans
    flat, spans = flatten(params, "grad")
    ensure_flat_grads._reg = (key, flat, spans)
    return flat, spans
