"""BNRefine (seg_oprs.py:143-162): defined by the reference for DFN but not instantiated by any shipped network, so it is
not covered by the DFN step test. Same structure as RefineResidual without the 1x1 (validated on hardware); first
green run: round-1 driver GPUTEST (x-passed); strict since round 2."""
import pytest
import torch

from util import norm_err
from test_gpu_bisenet import _sd_of, _prep, _rand, BN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("has_relu", [True, False])
def test_bn_refine_teacher_forced(cuda, has_relu):
    from torchseg_b200 import ops
    from torchseg_b200.seg_opr.seg_oprs import BNRefine
    from oracle import torch_ref as tr
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(41)
    mod = BNRefine(128, 128, 3, has_bias=False, has_relu=has_relu, norm_layer=BN)
    sd = _sd_of(mod)
    _prep(mod, cuda)
    x = _rand((8, 128, 24, 32), g, relu=True)
    xr = x.clone().requires_grad_(True)
    tr.set_bf16_emulation(True)
    try:
        yr = tr.bn_refine(xr, sd, "m", has_relu, 1e-5, 0.1, True)
        gy = _rand(tuple(yr.shape), g)
        yr.backward(gy)
    finally:
        tr.set_bf16_emulation(False)
    xd = ops.to_nhwc(x.to(cuda)).requires_grad_(True)
    yd = mod(xd)
    yd.backward(ops.to_nhwc(gy.to(cuda)))
    assert norm_err(yd, yr) < 1e-2
    assert norm_err(xd.grad, xr.grad) < 2e-2
    for n, p in mod.named_parameters():
        assert norm_err(p.grad, sd["m." + n].grad) < 2e-2, n
