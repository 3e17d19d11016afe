"""diagnostic: per-tensor forward/backward comparison of the B200 BiSeNet against the fp32 oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F
from util import norm_err, make_labels
import torchseg_b200
from torchseg_b200 import ops
from torchseg_b200.networks import BiSeNet
from torchseg_b200.seg_opr.loss_opr import ProbOhemCrossEntropy2d
from torchseg_b200.seg_opr.seg_oprs import upsample_bilinear
from torchseg_b200.utils.init_func import init_weight
from oracle import torch_ref as tr

cuda = torch.device("cuda:0")
torch.manual_seed(0)
N, H, W = int(os.environ.get("DN", 8)), int(os.environ.get("DH", 128)), int(os.environ.get("DH", 128))
mk = N * H * W // 16
crit = ProbOhemCrossEntropy2d(255, thresh=0.7, min_kept=mk)
model = BiSeNet(19, True, crit, None, torch.nn.BatchNorm2d)
init_weight(model.business_layer, torch.nn.init.kaiming_normal_, torch.nn.BatchNorm2d, 1e-5, 0.1, mode='fan_in', nonlinearity='relu')
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
for k in sd:
    if sd[k].is_floating_point() and "running" not in k:
        sd[k].requires_grad_(True)
model.to(cuda); torchseg_b200.prepare_model(model); model.train()
g = torch.Generator().manual_seed(1)
x = torch.randn(N, 3, H, W, generator=g); y = make_labels(N, H, W, 19, 255, g)

# ---- oracle with retained intermediates
kw = dict(eps=1e-5, momentum=0.1, training=True, stats=None)
R = {}
def keep(name, t):
    t.retain_grad(); R[name] = t; return t
tr.set_bf16_emulation(os.environ.get("EMU", "1") == "1")
sp = tr.conv_bn_relu(tr.q(x), sd, "spatial_path.conv_7x7", 2, 3, **kw)
sp = tr.conv_bn_relu(sp, sd, "spatial_path.conv_3x3_1", 2, 1, **kw)
sp = tr.conv_bn_relu(sp, sd, "spatial_path.conv_3x3_2", 2, 1, **kw)
sp = keep("spatial_out", tr.conv_bn_relu(sp, sd, "spatial_path.conv_1x1", 1, 0, **kw))
blocks = tr.resnet18(x, sd, "context_path", 1e-5, 0.1, True, None)
for i, b in enumerate(blocks): keep("block%d" % i, b)
blocks = blocks[::-1]
gc = tr.q(F.adaptive_avg_pool2d(blocks[0], 1))
gc = keep("gc", tr.conv_bn_relu(gc, sd, "global_context.1", 1, 0, **kw))
last = tr.q(F.interpolate(gc, size=blocks[0].shape[2:], mode="bilinear", align_corners=True))
po = []
for i in range(2):
    fm = tr.attention_refinement(blocks[i], sd, "arms.%d" % i, 1e-5, 0.1, True, None)
    fm = keep("arm%d" % i, tr.q(fm + last))
    last = tr.q(F.interpolate(fm, size=blocks[i + 1].shape[2:], mode="bilinear", align_corners=True))
    last = keep("refine%d" % i, tr.conv_bn_relu(last, sd, "refines.%d" % i, 1, 1, **kw))
    po.append(last)
po.append(keep("ffm", tr.feature_fusion(sp, last, sd, "ffm", 1e-5, 0.1, True, None)))
los = [keep("lo%d" % i, tr.bisenet_head_logits(po[i], sd, "heads.%d" % i, 1e-5, 0.1, True, None)) for i in range(3)]
losses = [tr.ohem_ce(F.interpolate(l, scale_factor=s, mode="bilinear", align_corners=True), y, 255, 0.7, mk) for l, s in zip(los, (16, 8, 8))]
(losses[2] + losses[0] + losses[1]).backward()

# ---- device with retained intermediates
D = {}
def keepd(name, t):
    t.retain_grad(); D[name] = t; return t
xd, yd = x.to(cuda), y.to(cuda)
spd = keepd("spatial_out", model.spatial_path(xd))
cb = model.context_path(xd)
for i, b in enumerate(cb): keepd("block%d" % i, b)
cb.reverse()
gcd = ops.AdaptiveAvgPoolFn.apply(cb[0], 1)
gcd = keepd("gc", model.global_context[1](gcd))
lastd = upsample_bilinear(gcd, cb[0].shape[2:])
pod = []
for i in range(2):
    fm = keepd("arm%d" % i, model.arms[i](cb[i], add=lastd))
    lastd = upsample_bilinear(fm, cb[i + 1].shape[2:])
    lastd = keepd("refine%d" % i, model.refines[i](lastd))
    pod.append(lastd)
pod.append(keepd("ffm", model.ffm(spd, lastd)))
lod = [keepd("lo%d" % i, model.heads[i].lowres_logits(pod[i])) for i in range(3)]
ld = [crit.forward_lowres(l, yd, 19) for l in lod]
(ld[2] + ld[0] + ld[1]).backward()
torch.cuda.synchronize()
print("losses ref", [float(l) for l in losses], "dev", [float(l) for l in ld])
for k in R:
    print("%-12s fwd %.4f  bwd %.4f   |g_ref| %.3e |g_dev| %.3e" % (k, norm_err(D[k], R[k]), norm_err(D[k].grad, R[k].grad),
          float(R[k].grad.norm()), float(D[k].grad.float().norm())))
print("---- params")
for n, p in model.named_parameters():
    print("%-50s %.4f" % (n, norm_err(p.grad, sd[n].grad)))
