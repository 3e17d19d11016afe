"""Parity of the DFN blocks and of the whole DFN-R101 training step against the oracle
(oracle/torch_ref.py, pinned to the live reference by tests/test_cpu_oracle.py::test_torch_oracle_dfn_vs_golden).

Same two layers of evidence as tests/test_gpu_bisenet.py: teacher-forced blocks against the bf16-storage-emulating
oracle (1e-2), then the whole step (loss 1e-2, classifier gradient direction, gradient magnitudes everywhere).
The ragged 21 / 171 / 9-channel layers run zero-padded to 64-channel multiples; the tests also check that the pad
channels stay exactly zero."""
import pytest
import torch

from util import norm_err, bf16_round
from test_gpu_bisenet import _sd_of, _prep, _rand, BN

pytestmark = pytest.mark.gpu


def _pad_c(t, Cp):
    if t.shape[1] == Cp:
        return t
    out = torch.zeros((t.shape[0], Cp) + tuple(t.shape[2:]), dtype=t.dtype)
    out[:, :t.shape[1]] = t
    return out


def _check_params(mod, sd, tol=2e-2):
    for n, p in mod.named_parameters():
        ref = sd["m." + n].grad
        assert p.grad is not None, n
        e = norm_err(p.grad, ref)
        assert e < tol, "param grad %s: %g" % (n, e)


@pytest.mark.parametrize("cfg", [(256, 21, True), (21, 21, True), (512, 171, False), (21, 9, False), (512, 512, True)],
                         ids=["256to21", "21to21", "512to171", "21to9", "512to512"])
def test_refine_residual_teacher_forced(cuda, cfg):
    """RefineResidual (seg_oprs.py:165-188) incl. the ragged channel counts of the border network / DFNHead"""
    from torchseg_b200 import ops
    from torchseg_b200.seg_opr.seg_oprs import RefineResidual
    from oracle import torch_ref as tr
    cin, cout, has_relu = cfg
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(11)
    mod = RefineResidual(cin, cout, 3, has_bias=False, has_relu=has_relu, norm_layer=BN)
    sd = _sd_of(mod)
    _prep(mod, cuda)
    x = _rand((8, cin, 24, 32), g, relu=True)
    xr = x.clone().requires_grad_(True)
    tr.set_bf16_emulation(True)
    try:
        yr = tr.refine_residual(xr, sd, "m", has_relu, 1e-5, 0.1, True)
        gy = _rand(tuple(yr.shape), g)
        yr.backward(gy)
    finally:
        tr.set_bf16_emulation(False)
    cin_p, cout_p = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    xd = ops.to_nhwc(_pad_c(x, cin_p).to(cuda)).requires_grad_(True)
    yd = mod(xd)
    assert yd.shape[1] == cout_p
    yd.backward(ops.to_nhwc(_pad_c(gy, cout_p).to(cuda)))
    assert norm_err(yd[:, :cout], yr) < 1e-2, norm_err(yd[:, :cout], yr)
    if cout_p != cout:
        assert float(yd[:, cout:].float().abs().max()) == 0.0, "pad channels must stay exactly zero"
    # dx sums two bf16 gradient streams (skip + 3x3 branch) that the oracle keeps in fp32: 2e-2
    assert norm_err(xd.grad[:, :cin], xr.grad) < 2e-2, norm_err(xd.grad[:, :cin], xr.grad)
    if cin_p != cin:
        assert float(xd.grad[:, cin:].float().abs().max()) == 0.0
    _check_params(mod, sd)
    # BN running statistics of the ragged layer are written back un-padded
    assert mod.cbr.bn.running_mean.shape[0] == cout and float(mod.cbr.bn.running_mean.abs().sum()) > 0


def test_channel_attention_teacher_forced(cuda):
    """ChannelAttention + SELayer (seg_oprs.py:110-140): x1 * sigmoid(fc(GAP(cat[x1,x2]))) + x2, nn.Linear on the conv path"""
    from torchseg_b200 import ops
    from torchseg_b200.seg_opr.seg_oprs import ChannelAttention
    from oracle import torch_ref as tr
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(12)
    mod = ChannelAttention(1024, 512, 1)
    sd = _sd_of(mod)
    _prep(mod, cuda)
    x1 = _rand((4, 512, 16, 24), g, relu=True)
    x2 = _rand((4, 512, 16, 24), g, relu=True)
    r1, r2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    tr.set_bf16_emulation(True)
    try:
        yr = tr.channel_attention(r1, r2, sd, "m")
        gy = _rand(tuple(yr.shape), g)
        yr.backward(gy)
    finally:
        tr.set_bf16_emulation(False)
    d1 = ops.to_nhwc(x1.to(cuda)).requires_grad_(True)
    d2 = ops.to_nhwc(x2.to(cuda)).requires_grad_(True)
    yd = mod(d1, d2)
    yd.backward(ops.to_nhwc(gy.to(cuda)))
    assert norm_err(yd, yr) < 1e-2
    assert norm_err(d1.grad, r1.grad) < 1e-2 and norm_err(d2.grad, r2.grad) < 1e-2
    _check_params(mod, sd, tol=3e-2)


@pytest.mark.parametrize("cfg", [(512, 19, 8), (21, 1, 4)], ids=["smooth", "border"])
def test_dfn_head_teacher_forced(cuda, cfg):
    """DFNHead (dfn network.py:157-172): RRB(in → 9·out) → 1x1 (+bias) → bilinear x scale, NCHW fp32 logits"""
    import torch.nn.functional as F
    from torchseg_b200 import ops
    from torchseg_b200.networks.dfn import DFNHead
    from oracle import torch_ref as tr
    cin, cout, scale = cfg
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(13)
    mod = DFNHead(cin, cout, scale, norm_layer=BN)
    sd = _sd_of(mod)
    _prep(mod, cuda)
    x = _rand((4, cin, 16, 24), g, relu=True)
    xr = x.clone().requires_grad_(True)
    tr.set_bf16_emulation(True)
    try:
        lo = tr.dfn_head_logits(xr, sd, "m", 1e-5, 0.1, True)
        yr = F.interpolate(lo, scale_factor=scale, mode="bilinear", align_corners=True)
        gy = torch.randn(tuple(yr.shape), generator=g)
        yr.backward(gy)
    finally:
        tr.set_bf16_emulation(False)
    cin_p = (cin + 63) // 64 * 64
    xd = ops.to_nhwc(_pad_c(x, cin_p).to(cuda)).requires_grad_(True)
    yd = mod(xd)
    assert tuple(yd.shape) == tuple(yr.shape) and yd.dtype == torch.float32
    yd.backward(gy.to(cuda))
    assert norm_err(yd, yr) < 1e-2, norm_err(yd, yr)
    assert norm_err(xd.grad[:, :cin], xr.grad) < 1e-2
    _check_params(mod, sd)


def test_focal_loss_on_border_logits(cuda):
    """SigmoidFocalLoss (loss_opr.py:14-45, a13) on [B,1,H,W] logits vs {0,1,255} labels, value and gradient"""
    from torchseg_b200.seg_opr.loss_opr import SigmoidFocalLoss
    from oracle import torch_ref as tr
    g = torch.Generator().manual_seed(14)
    pred = torch.randn(3, 1, 40, 56, generator=g) * 2
    tgt = (torch.rand(3, 40, 56, generator=g) < 0.2).long()
    tgt[:, :4] = 255
    pr = pred.clone().requires_grad_(True)
    lr = tr.sigmoid_focal(pr, tgt, 255)
    lr.backward()
    pd = pred.to(cuda).requires_grad_(True)
    ld = SigmoidFocalLoss(255)(pd, tgt.to(cuda))
    ld.backward()
    assert abs(ld.item() - lr.item()) < 1e-5 * abs(lr.item())
    assert norm_err(pd.grad, pr.grad) < 1e-4


def test_dfn_r101_step_matches_oracle(cuda):
    """DFN-R101_v1c (SURVEY C4 family at a small spatial size): smooth network with fused-upsample CE on 4 heads,
    border network with focal loss on 4 heads. Loss vs the fp32 oracle; gradient direction at the well-conditioned
    classifiers; gradient magnitude on every conv / linear weight."""
    import torchseg_b200
    from torchseg_b200.networks import DFN
    from torchseg_b200.seg_opr.loss_opr import SigmoidFocalLoss
    from oracle import torch_ref
    torch.manual_seed(5)
    N, HW = 8, 128
    m = DFN(19, torch.nn.CrossEntropyLoss(ignore_index=255), SigmoidFocalLoss(255), 0.1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(0, 19, (N, HW, HW), generator=g)
    y[:, :12] = 255
    e = (torch.rand(N, HW, HW, generator=g) < 0.15).long()
    e[:, :12] = 255
    loss_ref, _ = torch_ref.dfn_loss(x, y, e, sd)
    loss_ref.backward()
    m.to(cuda)
    torchseg_b200.prepare_model(m)
    m.train()
    loss = m(x.to(cuda), y.to(cuda), e.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 1e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    P = dict(m.named_parameters())
    for n in ("smooth_heads.3.conv.weight", "smooth_heads.2.conv.weight", "border_heads.3.conv.weight"):
        a, b = P[n].grad.float().cpu().reshape(-1), sd[n].grad.reshape(-1)
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
        assert cos > 0.85, (n, cos)    # 0.889 observed once on the deepest head: 8 images x 16x16 positions only
    checked, bad = 0, []
    for n, p in P.items():
        if p.dim() in (2, 4):
            if sd[n].grad is None:   # border_aft_rrbs.0 is constructed but never called (dfn network.py:128-135)
                assert n.startswith("border_aft_rrbs.0.") and (p.grad is None or float(p.grad.abs().max()) == 0.0), n
                continue
            assert p.grad is not None, n
            ratio = float(p.grad.float().norm().cpu() / sd[n].grad.norm().clamp_min(1e-30))
            if not 0.66 < ratio < 1.5:   # 9- and 21-channel border layers: few elements, noisy norms (0.74 observed)
                bad.append((n, round(ratio, 3)))
            checked += 1
    assert checked >= 180 and not bad, bad[:12]
