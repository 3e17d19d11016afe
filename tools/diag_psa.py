"""diagnostic: where does the PSA module's input-gradient error come from"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
from util import norm_err, bf16_round
from torchseg_b200 import ops
import torchseg_b200
cuda = torch.device("cuda")
g = torch.Generator().manual_seed(1)

def rnd(shape, relu=False, s=1.0):
    t = torch.randn(*shape, generator=g) * s
    return bf16_round(torch.relu(t) if relu else t)

# 1. PSABmmFn alone, c = 512, peaky attention
for c, sc in ((128, 2.0), (512, 2.0), (512, 5.0)):
    b, h, w = 2, 60, 60
    L = h * w
    x, att, gy = rnd((b, c, h, w), True), rnd((b, L, h, w), s=sc), rnd((b, c, h, w))
    xr, ar = x.clone().requires_grad_(True), att.clone().requires_grad_(True)
    yr = torch.bmm(xr.view(b, c, -1), torch.softmax(ar.view(b, L, -1), dim=1)).view(b, c, h, w)
    yr.backward(gy)
    xd = ops.to_nhwc(x.to(cuda)).requires_grad_(True)
    ad = ops.to_nhwc(att.to(cuda), dtype=torch.float32).requires_grad_(True)
    yd = ops.PSABmmFn.apply(xd, ad)
    yd.backward(ops.to_nhwc(gy.to(cuda)))
    print("bmm c=%d scale=%g: y %.4f dx %.4f datt %.4f" % (c, sc, norm_err(yd, yr), norm_err(xd.grad, xr.grad), norm_err(ad.grad, ar.grad)))

# 2. ConvFn 512 -> 3600 (fp32 out), bf16 dy with channel stride 3648
conv = torch.nn.Conv2d(512, 3600, 1, bias=False)
x = rnd((2, 512, 60, 60), True)
gy = rnd((2, 3600, 60, 60))
xr = x.clone().requires_grad_(True)
wr = bf16_round(conv.weight.detach()).requires_grad_(True)
yr = F.conv2d(xr, wr)
yr.backward(gy)
conv.to(cuda); torchseg_b200.prepare_model(conv)
from torchseg_b200.seg_opr.seg_oprs import conv_plain
xd = ops.to_nhwc(x.to(cuda)).requires_grad_(True)
yd = conv_plain(xd, conv, out_f32=True)
gyd = ops.nhwc_zeros(2, 3600, 60, 60, device=cuda, cs=3648)
gyd.copy_(gy.to(cuda))
yd.backward(gyd)
print("conv512->3600: y %.4f dx %.4f dw %.4f" % (norm_err(yd, yr), norm_err(xd.grad, xr.grad), norm_err(conv.weight.grad, wr.grad)))

# 3. the module, all errors
from torchseg_b200.networks.psanet import PointwiseSpatialAttention
from oracle import torch_ref as tr
torch.manual_seed(0)
mod = PointwiseSpatialAttention('psa', 150, 2048, norm_layer=torch.nn.BatchNorm2d)
mod.conv6[1].p = 0.0
with torch.no_grad():
    mod.collect_attention[1].conv.weight.mul_(12.0)
    mod.distribute_attention[1].conv.weight.mul_(12.0)
sd = {"m." + k: v.detach().clone() for k, v in mod.state_dict().items()}
for k, v in sd.items():
    if v.is_floating_point() and "running" not in k:
        v.requires_grad_(True)
mod.to(cuda); torchseg_b200.prepare_model(mod); mod.train()
x = rnd((2, 2048, 60, 60), True)
xr = x.clone().requires_grad_(True)
for emu in (True, False):
    for v in sd.values():
        v.grad = None
    xr.grad = None
    tr.set_bf16_emulation(emu)
    yr = tr.psa_logits(xr, sd, "m", 1e-5, 0.1, True)
    if emu:
        gy = torch.randn(tuple(yr.shape), generator=g)
    yr.backward(gy)
    tr.set_bf16_emulation(False)
    if emu:
        ref_emu = (yr.detach().clone(), xr.grad.clone(), {k: v.grad.clone() for k, v in sd.items() if v.grad is not None})
    else:
        print("oracle fp32 vs oracle bf16-emulation: y %.4f dx %.4f" % (norm_err(ref_emu[0], yr), norm_err(ref_emu[1], xr.grad)))
        for k, v in sd.items():
            if v.grad is not None and v.dim() == 4:
                print("   emu-vs-fp32 %-50s %.4f" % (k, norm_err(ref_emu[2][k], v.grad)))
yr, xg, pg = ref_emu
xd = ops.to_nhwc(x.to(cuda)).requires_grad_(True)
yd = mod(xd)
yd.backward(gy.to(cuda))
print("module: y %.4f dx %.4f" % (norm_err(yd, yr), norm_err(xd.grad, xg)))
for n, p in mod.named_parameters():
    print("   %-50s %.4f" % (n, norm_err(p.grad, pg["m." + n])))
