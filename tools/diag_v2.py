"""A/B the persistent conv kernel switches on row-tile cases: base_offset on/off, rows on/off, resident on/off"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from torchseg_b200 import ops, _lib
cuda = torch.device("cuda:0")
bf = lambda t: t.to(torch.bfloat16).float()
def run(case, tag):
    N, C, H, W, K, R, st, pad, dil = case
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(N, C, H, W, generator=g)); w = bf(torch.randn(K, C, R, R, generator=g) / (C * R * R) ** 0.5)
    ref = F.conv2d(x, w, None, st, pad, dil)
    xd = ops.to_nhwc(x.to(cuda)); wb = w.to(cuda).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    try:
        y = ops.conv_fprop(xd, wb, K, R, st, pad, dil, out_dtype=torch.float32)
        torch.cuda.synchronize()
        e = float((y.cpu() - ref).abs().max() / ref.abs().max())
    except Exception as ex:
        e = "EXC %s" % str(ex)[:80]
    print(tag, case, e)
cases = [(1, 64, 6, 160, 64, 3, 1, 1, 1), (1, 64, 7, 260, 128, 3, 2, 1, 1), (2, 64, 32, 32, 64, 3, 1, 1, 1)]
for bo in (1, 0):
    _lib.call("tsb_debug_set", 1, bo)
    for c in cases:
        run(c, "base_offset=%d" % bo)
_lib.call("tsb_debug_set", 1, 1)
for key, name in ((2, "rows"), (3, "resident")):
    _lib.call("tsb_debug_set", key, 0)
    for c in cases:
        run(c, "%s=0" % name)
    _lib.call("tsb_debug_set", key, 1)
