import torch


def rel_err(a, b):
    """max |a-b| / max|b| (scale-relative, robust for bf16 comparisons)"""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    denom = b.abs().max().clamp_min(1e-12)
    return float((a - b).abs().max() / denom)


def norm_err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def make_labels(N, H, W, C, ignore, gen, ignore_frac=0.1):
    lab = torch.randint(0, C, (N, H, W), generator=gen, dtype=torch.int64)
    k = max(1, int(H * ignore_frac))
    lab[:, :k, :] = ignore  # deterministic ~10 % ignore band
    return lab
