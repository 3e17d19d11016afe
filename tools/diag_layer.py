"""diagnostic: single-layer numerics of conv (fp32 out) / BN stats / BN apply against fp32 torch on bf16-exact inputs"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torchseg_b200 import ops

cuda = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
bf = lambda t: t.to(torch.bfloat16).float()
N, C, H, W, K = 8, 64, 64, 64, 64
x = bf(torch.relu(torch.randn(N, C, H, W, generator=g)))
w = bf(torch.randn(K, C, 3, 3, generator=g) * (2.0 / (C * 9)) ** 0.5)
ref = F.conv2d(x.double(), w.double(), None, 1, 1).float()           # exact reference
ref32 = F.conv2d(x, w, None, 1, 1)
xd = ops.to_nhwc(x.to(cuda))
wb = w.to(cuda).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
y32 = ops.conv_fprop(xd, wb, K, 3, 1, 1, 1, out_dtype=torch.float32).cpu()
cud = F.conv2d(x.to(cuda), w.to(cuda), None, 1, 1).cpu()
n = lambda a, b: float((a - b).norm() / b.norm())
print("conv fp32-out: mine vs exact %.3e | cpu fp32 vs exact %.3e | cudnn fp32 vs exact %.3e" % (n(y32, ref), n(ref32, ref), n(cud, ref)))
stats = torch.zeros(2, K, device=cuda)
yb = ops.conv_fprop(xd, wb, K, 3, 1, 1, 1, stats=stats).float().cpu()
qref = bf(ref)
flip = (yb != qref).float().mean().item()
print("conv bf16-out: flipped fraction vs q(exact) %.3e, norm err %.3e" % (flip, n(yb, qref)))
# BN statistics
m_ref = qref.double().mean(dim=(0, 2, 3)); v_ref = qref.double().var(dim=(0, 2, 3), unbiased=False)
cnt = N * H * W
m = stats[0].cpu().double() / cnt; v = stats[1].cpu().double() / cnt - m * m
print("stats: mean rel err %.3e  var rel err %.3e" % (float(((m - m_ref).abs() / v_ref.sqrt()).max()), float(((v - v_ref).abs() / v_ref).max())))
# BN apply
gamma = torch.rand(K, generator=g) + 0.5; beta = torch.randn(K, generator=g)
aux = torch.empty(4, K, device=cuda)
ops.call("tsb_bn_finalize", ops.ptr(stats[0]), ops.ptr(stats[1]), float(cnt), K, ops.ptr(gamma.to(cuda)), ops.ptr(beta.to(cuda)),
         1e-5, 0.1, ops.ptr(aux[0]), ops.ptr(aux[1]), ops.ptr(aux[2]), ops.ptr(aux[3]), None, None, ops.stream())
ybd = ops.to_nhwc(qref.to(cuda))
out = ops.nhwc_empty(N, K, H, W)
ops.call("tsb_bn_apply", ops.ptr(ybd), K, ops.ptr(aux[2]), ops.ptr(aux[3]), None, 0, 1, ops.ptr(out), K, N * H * W, K, ops.stream())
bn_ref = F.relu(F.batch_norm(qref.double(), None, None, gamma.double(), beta.double(), True, 0.1, 1e-5)).float()
outc = out.float().cpu()
print("bn apply: flipped fraction vs q(exact) %.3e, norm err %.3e ; pre-round check invstd rel err %.3e" % (
    (outc != bf(bn_ref)).float().mean().item(), n(outc, bf(bn_ref)),
    float(((aux[1].cpu().double() - 1 / (v_ref + 1e-5).sqrt()).abs() * (v_ref + 1e-5).sqrt()).max())))
# module-level: ConvBnRelu vs emulated oracle
from oracle import torch_ref as tr
from torchseg_b200.seg_opr.seg_oprs import ConvBnRelu
import torchseg_b200
torch.manual_seed(1)
mod = ConvBnRelu(C, K, 3, 1, 1)
sd = {"m." + k: v.detach().clone() for k, v in mod.state_dict().items()}
mod.to(cuda); torchseg_b200.prepare_model(mod); mod.train()
tr.set_bf16_emulation(True)
yr = tr.conv_bn_relu(x, sd, "m", 1, 1)
yd = mod(x.to(cuda)).float().cpu()
print("ConvBnRelu module vs emulated oracle: flipped %.3e norm err %.3e" % ((yd != yr).float().mean().item(), n(yd, yr)))
