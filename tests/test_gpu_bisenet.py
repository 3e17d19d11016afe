"""End-to-end parity of the B200 BiSeNet-R18 training step against the fp32 oracle (oracle/torch_ref.py, itself
pinned to the live reference): loss, low-resolution logits, parameter gradients. bf16 tolerance."""
import pytest
import torch

from util import rel_err, norm_err, make_labels

pytestmark = pytest.mark.gpu


def _build(cuda, seed=0):
    import torchseg_b200
    from torchseg_b200.networks import BiSeNet
    from torchseg_b200.seg_opr.loss_opr import ProbOhemCrossEntropy2d
    from torchseg_b200.utils.init_func import init_weight
    torch.manual_seed(seed)
    N, H, W = 2, 256, 256
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=N * H * W // 16, use_weight=False)
    model = BiSeNet(19, True, crit, None, torch.nn.BatchNorm2d)
    init_weight(model.business_layer, torch.nn.init.kaiming_normal_, torch.nn.BatchNorm2d, 1e-5, 0.1, mode='fan_in',
                nonlinearity='relu')
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.to(cuda)
    torchseg_b200.prepare_model(model)
    model.train()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(N, 3, H, W, generator=g)
    y = make_labels(N, H, W, 19, 255, g)
    return model, sd, x, y, N * H * W // 16


def test_bisenet_step_matches_oracle(cuda):
    from oracle import torch_ref
    model, sd, x, y, min_kept = _build(cuda)
    for k in sd:
        if sd[k].is_floating_point() and "running" not in k:
            sd[k].requires_grad_(True)
    stats = {}
    loss_ref, lo_ref = torch_ref.bisenet_r18_loss(x, y, sd, min_kept, stats=stats)
    loss_ref.backward()
    loss = model(x.to(cuda), y.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    worst = 0.0
    bad = []
    for n, p in model.named_parameters():
        ref = sd[n].grad
        e = norm_err(p.grad, ref)
        worst = max(worst, e)
        if e > 0.1:
            bad.append((n, e))
    assert not bad, "gradient mismatch (norm-relative > 0.1): %s" % bad[:10]
    # BN running statistics follow the reference update (momentum 0.1, unbiased variance)
    msd = model.state_dict()
    for k, v in stats.items():
        assert rel_err(msd[k], v) < 3e-2, k


def test_bisenet_train_steps_decrease_loss(cuda):
    """three optimiser steps with the fused flat SGD: loss decreases, parameters stay finite"""
    from torchseg_b200 import optim
    from torchseg_b200.utils.init_func import group_weight
    model, sd, x, y, _ = _build(cuda, seed=3)
    groups = []
    groups = group_weight(groups, model.context_path, torch.nn.BatchNorm2d, 1e-2)
    for m in (model.spatial_path, model.global_context, model.arms, model.refines, model.heads, model.ffm):
        groups = group_weight(groups, m, torch.nn.BatchNorm2d, 1e-1)
    opt = optim.SGD(groups, lr=1e-2, momentum=0.9, weight_decay=5e-4)
    xs, ys = x.to(cuda), y.to(cuda)
    losses = []
    for it in range(4):
        opt.zero_grad()
        loss = model(xs, ys)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(l == l for l in losses)
    assert losses[-1] < losses[0], losses
