"""Parity of the PSANet head against the oracle (oracle/torch_ref.py, pinned to the live reference
by tests/test_cpu_oracle.py::test_torch_oracle_psanet_vs_golden): the channel soft-max + bmm contraction kernels
alone, the PointwiseSpatialAttention module teacher-forced, and the whole PSANet-R101 step."""
import pytest
import torch

from util import norm_err, rel_err, bf16_round
from test_gpu_bisenet import _sd_of, _prep, _rand, BN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(8, 16), (60, 60)], ids=["128pos", "3600pos"])
def test_psa_bmm_matches_torch(cuda, hw):
    """bmm(x.view(b,c,-1), softmax(att.view(b,L,-1), dim=1)) and its gradients vs fp32 torch on bf16-exact inputs"""
    from torchseg_b200 import ops
    h, w = hw
    L = h * w
    b, c = 2, 128
    g = torch.Generator().manual_seed(21)
    x = _rand((b, c, h, w), g, relu=True)
    att = bf16_round(torch.randn(b, L, h, w, generator=g) * 2)
    gy = _rand((b, c, h, w), g)
    xr, ar = x.clone().requires_grad_(True), att.clone().requires_grad_(True)
    S = torch.softmax(ar.view(b, L, -1), dim=1)
    yr = torch.bmm(xr.view(b, c, -1), S).view(b, c, h, w)
    yr.backward(gy)
    xd = ops.to_nhwc(x.to(cuda)).requires_grad_(True)
    ad = ops.to_nhwc(att.to(cuda)).requires_grad_(True)
    yd = ops.PSABmmFn.apply(xd, ad)
    yd.backward(ops.to_nhwc(gy.to(cuda)))
    assert norm_err(yd, yr) < 1e-2, norm_err(yd, yr)
    assert norm_err(xd.grad, xr.grad) < 1e-2, norm_err(xd.grad, xr.grad)
    assert norm_err(ad.grad, ar.grad) < 2e-2, norm_err(ad.grad, ar.grad)


def test_softmax_rows_properties(cuda):
    """rows sum to 1 (bf16 resolution), pad channels are exact zeros, shift invariance"""
    from torchseg_b200 import ops
    g = torch.Generator().manual_seed(22)
    a = bf16_round(torch.randn(1, 3600, 4, 8, generator=g) * 3)
    ad = ops.to_nhwc(a.to(cuda))
    S = ops.nhwc_empty(1, 3648, 4, 8, device=cuda)
    ops.call("tsb_softmax_rows_fwd", ops.ptr(ad), ops.BF16, 3600, ops.ptr(S), 3648, 32, 3600, 3648, ops.stream())
    Sf = S.float()
    assert float(Sf[:, 3600:].abs().max()) == 0.0
    assert float((Sf.sum(1) - 1).abs().max()) < 5e-3
    ref = torch.softmax(a, dim=1)
    assert rel_err(Sf[:, :3600], ref) < 1e-2
    a2 = ops.to_nhwc((a + 4.0).to(cuda))     # bf16(a + 4) is not exactly a + 4 for every element: compare loosely
    S2 = ops.nhwc_empty(1, 3648, 4, 8, device=cuda)
    ops.call("tsb_softmax_rows_fwd", ops.ptr(a2), ops.BF16, 3600, ops.ptr(S2), 3648, 32, 3600, 3648, ops.stream())
    assert norm_err(S2[:, :3600], torch.softmax(bf16_round(a + 4.0), dim=1)) < 1e-2


def test_pointwise_spatial_attention_teacher_forced(cuda):
    """PointwiseSpatialAttention (psanet network.py:75-144) on a [2,2048,60,60] block, Dropout2d off"""
    from torchseg_b200 import ops
    from torchseg_b200.networks.psanet import PointwiseSpatialAttention
    from oracle import torch_ref as tr
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(23)
    mod = PointwiseSpatialAttention('psa', 150, 2048, norm_layer=BN)
    mod.conv6[1].p = 0.0
    # With the default init the attention logits have std ~0.4, the soft-max over 3600 positions is almost uniform and
    # both branches return a spatially (almost) constant map; the BN of `proj` then rescales the residual variation —
    # which is at the level of the bf16 rounding of that constant — to unit variance, an ill-conditioned comparison
    # (4.5e-2 observed, identical with fp32 logits). Sharpen the attention (std ~5: a trained head is peaky) so the
    # test measures the kernels and not that amplification.
    with torch.no_grad():
        mod.collect_attention[1].conv.weight.mul_(12.0)
        mod.distribute_attention[1].conv.weight.mul_(12.0)
    sd = _sd_of(mod)
    _prep(mod, cuda)
    x = _rand((2, 2048, 60, 60), g, relu=True)
    xr = x.clone().requires_grad_(True)
    gy = None
    runs = {}
    for emu in (True, False):
        for v in sd.values():
            v.grad = None
        xr.grad = None
        tr.set_bf16_emulation(emu)
        try:
            yr = tr.psa_logits(xr, sd, "m", 1e-5, 0.1, True)
            if gy is None:
                gy = torch.randn(tuple(yr.shape), generator=g)
            yr.backward(gy)
        finally:
            tr.set_bf16_emulation(False)
        runs[emu] = (yr.detach().clone(), xr.grad.clone(), {k: v.grad.clone() for k, v in sd.items() if v.grad is not None})
    (yr, xg, pg), (y32, xg32, pg32) = runs[True], runs[False]
    xd = ops.to_nhwc(x.to(cuda)).requires_grad_(True)
    yd = mod(xd)
    assert yd.dtype == torch.float32 and tuple(yd.shape) == tuple(yr.shape)
    yd.backward(gy.to(cuda))
    assert norm_err(yd, yr) < 1e-2, norm_err(yd, yr)
    # Backward: this head is ill-conditioned with respect to bf16 storage even when sharpened — the fp32 oracle and its
    # own bf16-storage emulation differ by ~15 % in dx and in every weight gradient below conv6 (tools/diag_psa.py;
    # the contraction kernels alone are at 2e-3..4e-3, test_psa_bmm_matches_torch). The CUDA path must sit well inside
    # that spread: closer to the bf16 model than half the distance between the two oracles (observed 5.7e-2 vs 14.6e-2).
    spread = norm_err(xg32, xg)
    e = norm_err(xd.grad, xg)
    assert e < max(2e-2, 0.5 * spread), (e, spread)
    for n, p in mod.named_parameters():
        ref, ref32 = pg["m." + n], pg32["m." + n]
        e = norm_err(p.grad, ref)
        assert e < max(3e-2, 0.6 * norm_err(ref32, ref), 0.6 * spread), "param grad %s: %g (oracle spread %g)" % (n, e, norm_err(ref32, ref))


def test_psanet_r101_step_matches_oracle(cuda):
    """PSANet-R101_v1c at the reference shape (480x480 → 60x60 = 3600 positions), 150 classes, batch 2"""
    import torchseg_b200
    from torchseg_b200.networks import PSANet
    from oracle import torch_ref
    torch.manual_seed(2)
    N, HW = 2, 480
    m = PSANet(150, torch.nn.CrossEntropyLoss(ignore_index=-1))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(-1, 150, (N, HW, HW), generator=g)
    torch.set_num_threads(min(32, torch.get_num_threads()))
    loss_ref, _ = torch_ref.psanet_loss(x, y, sd)
    loss_ref.backward()
    m.to(cuda)
    torchseg_b200.prepare_model(m)
    m.train()
    loss = m(x.to(cuda), y.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 1e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    P = dict(m.named_parameters())
    for n in ("psa_layer.conv6.2.weight", "aux_layer.2.weight"):
        a, b = P[n].grad.float().cpu().reshape(-1), sd[n].grad.reshape(-1)
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
        assert cos > 0.9, (n, cos)
    checked, bad = 0, []
    for n, p in P.items():
        if p.dim() == 4:
            ratio = float(p.grad.float().norm().cpu() / sd[n].grad.norm().clamp_min(1e-30))
            if not 0.75 < ratio < 1.33:
                bad.append((n, round(ratio, 3)))
            checked += 1
    assert checked >= 100 and not bad, bad[:12]
