"""Parity AT THE BASELINE SHAPE (BASELINE.json configs[1]: BiSeNet-R18, 16 x 1024 x 1024, 19 classes, OHEM): the whole
training step of the libtsb path against the oracle (oracle/torch_ref.py, bit-identical to the live reference modules)
run ON THE SAME GPU under stock torch.nn + cuDNN in strict fp32 (TF32 off) — SURVEY.md §8c "On the GPU box", north_star
"outputs match the reference PyTorch/cuDNN path".

The row-tile / PAIR / persistent conv paths, the band OHEM kernels and the multi-wave streaming kernels only run at this
size, so this is where they are checked against the oracle (the small-shape tests exercise other tile choices).
Tolerances: loss 1e-2 (north_star, bf16); gradients of the business layers (heads, FFM, ARMs, refines, global context,
spatial path) norm_err <= 2e-2 per conv / classifier weight; BN running statistics 1e-2; the OHEM kept count of every
head equal to the oracle's."""
import pytest
import torch
import torch.nn.functional as F

from util import norm_err, make_labels

pytestmark = pytest.mark.gpu
BN = torch.nn.BatchNorm2d


def _oracle_kept_counts(lo, labels, min_kept):
    from oracle import torch_ref
    out = []
    for l, s in zip(lo, (16, 8, 8)):
        up = F.interpolate(l.detach(), scale_factor=s, mode="bilinear", align_corners=True)
        _, valid, thr = torch_ref.ohem_ce(up, labels, 255, 0.7, min_kept, return_aux=True)
        out.append((int(valid.sum()), float(thr)))
        del up
    return out


def test_bisenet_full_size_step_matches_oracle_cuda(cuda):
    import torchseg_b200
    from torchseg_b200 import ops
    from torchseg_b200.networks import BiSeNet
    from torchseg_b200.seg_opr.loss_opr import ProbOhemCrossEntropy2d
    from torchseg_b200.utils.init_func import init_weight
    from oracle import torch_ref
    H = W = 1024
    free, _ = torch.cuda.mem_get_info()
    N = 16 if free > 120 * (1 << 30) else 8
    min_kept = N * H * W // 16                                   # train.py:48-49
    torch.manual_seed(12345)
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=min_kept, use_weight=False)
    crit.debug_states = []
    model = BiSeNet(19, True, crit, None, BN)
    init_weight(model.business_layer, torch.nn.init.kaiming_normal_, BN, 1e-5, 0.1, mode='fan_in', nonlinearity='relu')
    sd = {k: v.detach().clone().to(cuda) for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, 3, H, W, generator=g).to(cuda)
    y = make_labels(N, H, W, 19, 255, g).to(cuda)

    # ---- oracle on the GPU: stock torch.nn lowering, cuDNN, strict fp32
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True                        # train.py:35
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        stats = {}
        loss_ref, lo_ref = torch_ref.bisenet_r18_loss(x, y, sd, min_kept, stats=stats)
        loss_ref.backward()
        kept_ref = _oracle_kept_counts(lo_ref, y, min_kept)
        loss_ref_v = float(loss_ref)
        grads_ref = {k: v.grad.detach().cpu() for k, v in sd.items() if v.grad is not None}
        stats = {k: v.detach().cpu() for k, v in stats.items()}
        del loss_ref, lo_ref
        for v in sd.values():
            v.grad = None
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    # ---- the libtsb step
    model.to(cuda)
    torchseg_b200.prepare_model(model)
    model.train()
    loss = model(x, y)
    loss.backward()
    torch.cuda.synchronize()
    rel = abs(loss.item() - loss_ref_v) / abs(loss_ref_v)
    print("full-size N=%d: loss %.6f oracle %.6f rel %.2e" % (N, loss.item(), loss_ref_v, rel))
    assert rel < 1e-2, (loss.item(), loss_ref_v)

    # OHEM: kept count per head (aux0, aux1, main — the order the network calls the criterion in)
    assert len(crit.debug_states) == 3
    for i, (st, (kept, thr)) in enumerate(zip(crit.debug_states, kept_ref)):
        d = ops.ohem_state_dict(st)
        n_kept = d["kept"]
        print("head %d: kept %d (oracle %d), threshold %.6f (oracle %.6f)" % (i, n_kept, kept, d["T"] if d["active"] else 0.7, thr))
        assert n_kept == kept, (i, n_kept, kept)

    # gradients: business layers tight, whole network aligned
    business = ("spatial_path.", "global_context.", "arms.", "refines.", "heads.", "ffm.")
    worst, bad, cos_bad = [], [], []
    for n, p in model.named_parameters():
        a, b = p.grad.float().cpu(), grads_ref[n]
        e = norm_err(a, b)
        af, bf = a.reshape(-1), b.reshape(-1)
        cos = float(torch.dot(af, bf) / (af.norm() * bf.norm()).clamp_min(1e-30))
        worst.append((e, n))
        if n.startswith(business) and p.dim() == 4 and e > 2e-2:
            bad.append((n, round(e, 4)))
        if cos < 0.95:
            cos_bad.append((n, round(cos, 4)))
    worst.sort(reverse=True)
    print("worst gradient norm_err:", [(n, round(e, 4)) for e, n in worst[:8]])
    assert not bad, "business-layer gradient norm_err > 2e-2: %s" % bad[:10]
    assert not cos_bad, "gradient cosine < 0.95: %s" % cos_bad[:10]

    # BN running statistics (momentum 0.1, unbiased variance)
    msd = model.state_dict()
    for k, v in stats.items():
        e = float((msd[k].float().cpu() - v).abs().max() / v.abs().max().clamp_min(1e-6))
        assert e < 1e-2, (k, e)
