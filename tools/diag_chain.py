"""diagnostic: teacher-forced per-layer error of the spatial path + resnet stem/layer1 vs the emulated oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import torchseg_b200
from torchseg_b200 import ops
from torchseg_b200.networks import BiSeNet
from torchseg_b200.utils.init_func import init_weight
from oracle import torch_ref as tr

cuda = torch.device("cuda:0")
torch.manual_seed(0)
model = BiSeNet(19, True, None, None, torch.nn.BatchNorm2d)
init_weight(model.business_layer, torch.nn.init.kaiming_normal_, torch.nn.BatchNorm2d, 1e-5, 0.1, mode='fan_in', nonlinearity='relu')
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
model.to(cuda); torchseg_b200.prepare_model(model); model.train()
g = torch.Generator().manual_seed(1)
x = torch.randn(8, 3, 128, 128, generator=g)
tr.set_bf16_emulation(True)
n = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
fl = lambda a, b: float((a.float().cpu() != b).float().mean())
with torch.no_grad():
    sp = model.spatial_path
    names = [("conv_7x7", 2, 3), ("conv_3x3_1", 2, 1), ("conv_3x3_2", 2, 1), ("conv_1x1", 1, 0)]
    ref_in = tr.q(x)
    dev_chain = x.to(cuda)
    for nm, st, pd in names:
        ref_out = tr.conv_bn_relu(ref_in, sd, "spatial_path." + nm, st, pd)
        mod = getattr(sp, nm)
        forced = mod(ref_in.to(cuda) if nm != "conv_7x7" else x.to(cuda))
        dev_chain = mod(dev_chain)
        # raw conv compare
        raw_ref = tr.q(F.conv2d(ref_in, tr.qw(sd["spatial_path.%s.conv.weight" % nm]), None, st, pd))
        print("%-12s teacher-forced: flipped %.3e err %.3e | chained err %.3e | raw mean %.3e std %.3e" % (
            nm, fl(forced, ref_out), n(forced, ref_out), n(dev_chain, ref_out), float(raw_ref.mean()), float(raw_ref.std())))
        ref_in = ref_out
    # resnet stem + maxpool + layer1
    cp = model.context_path
    r = tr.q(F.conv2d(tr.q(x), tr.qw(sd["context_path.conv1.weight"]), None, 2, 3))
    r = tr.q(F.relu(tr._bn(r, sd, "context_path.bn1", 1e-5, 0.1, True)))
    from torchseg_b200.seg_opr.seg_oprs import conv_bn_act
    d = conv_bn_act(x.to(cuda), cp.conv1, cp.bn1, True)
    print("resnet stem: flipped %.3e err %.3e" % (fl(d, r), n(d, r)))
    r2 = F.max_pool2d(r, 3, 2, 1); d2 = ops.MaxPool3x3S2Fn.apply(ops.to_nhwc(r.to(cuda)))
    print("maxpool teacher-forced err %.3e" % n(d2, r2))
    rin = r2
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for bi in range(2):
            st = stride if bi == 0 else 1
            pre = "context_path.layer%d.%d" % (li, bi)
            rout = tr.basic_block(rin, sd, pre, st, bi == 0 and li > 1, 1e-5, 0.1, True)
            blk = getattr(cp, "layer%d" % li)[bi]
            dout = blk(ops.to_nhwc(rin.to(cuda)))
            print("%-24s teacher-forced: flipped %.3e err %.3e" % (pre, fl(dout, rout), n(dout, rout)))
            rin = rout
