"""B200-native building blocks with the reference's names, constructor signatures and child-module names
(/root/reference/furnace/seg_opr/seg_oprs.py). Parameters live in genuine nn.Conv2d / norm_layer children —
`group_weight`'s completeness assert (furnace/utils/init_func.py:52-53) and the checkpoint keys (SURVEY.md
App. D) depend on that — only `forward` is replaced by the fused libtsb path.
"""
import torch
import torch.nn as nn

from .. import ops


def _as_act(x):
    """boundary: accept a plain NCHW tensor (any dtype) and move it into the NHWC bf16 activation layout"""
    if x.dtype == torch.bfloat16 and (x.shape[1] == 1 or x.stride(1) == 1):
        return x
    return ops.to_nhwc(x)


_s2d_cache = {}


def _packed_image(img):
    """space-to-depth pack of the fp32 NCHW input image, shared by the two 7x7 stems of BiSeNet"""
    key = (img.data_ptr(), img._version, tuple(img.shape))
    hit = _s2d_cache.get("k")
    if hit is not None and hit[0] == key:
        return hit[1]
    packed = ops.pack_image_s2d(img.contiguous().float())
    _s2d_cache["k"] = (key, packed)
    return packed


def conv_bn_act(x, conv, bn, relu, residual=None):
    """fused conv → norm_layer(train/eval) → (+residual) → ReLU on the libtsb path"""
    ks = conv.kernel_size[0]
    assert conv.kernel_size[0] == conv.kernel_size[1] and conv.groups == 1 and conv.bias is None
    stem = (conv.in_channels == 3 and conv.stride[0] == 2 and conv.dilation[0] == 1 and
            (ks, conv.padding[0]) in ((7, 3), (3, 1)))
    if stem:
        xin = _packed_image(x) if x.shape[1] == 3 else x
    else:
        xin = _as_act(x)
    if residual is not None:
        residual = _as_act(residual)
    momentum = bn.momentum if bn.momentum is not None else 0.1
    return ops.ConvBNActFn.apply(xin, conv.weight, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var,
                                 conv.stride[0], conv.padding[0], conv.dilation[0], bool(relu), float(bn.eps),
                                 float(momentum), bool(bn.training), stem)


def conv_plain(x, conv, out_f32=False, ocs=None):
    return ops.ConvFn.apply(_as_act(x), conv.weight, conv.bias, conv.stride[0], conv.padding[0], conv.dilation[0],
                            bool(out_f32), ocs)


class ConvBnRelu(nn.Module):
    """seg_oprs.py:24-46 — same ctor, children named conv / bn / relu"""

    def __init__(self, in_planes, out_planes, ksize, stride, pad, dilation=1, groups=1, has_bn=True,
                 norm_layer=nn.BatchNorm2d, bn_eps=1e-5, has_relu=True, inplace=True, has_bias=False):
        super(ConvBnRelu, self).__init__()
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=ksize, stride=stride, padding=pad,
                              dilation=dilation, groups=groups, bias=has_bias)
        self.has_bn = has_bn
        if self.has_bn:
            self.bn = norm_layer(out_planes, eps=bn_eps)
        self.has_relu = has_relu
        if self.has_relu:
            self.relu = nn.ReLU(inplace=inplace)

    def forward(self, x):
        if self.has_bn and self.conv.bias is None:
            return conv_bn_act(x, self.conv, self.bn, self.has_relu)
        y = conv_plain(x, self.conv)
        if self.has_bn:
            raise NotImplementedError("ConvBnRelu with both bias and BN is not on the accelerated path")
        if self.has_relu:
            y = ReluFn.apply(y)
        return y


class ReluFn(torch.autograd.Function):
    """stand-alone ReLU on a small tensor (FFM squeeze branch [N,C,1,1], seg_oprs.py:224-226): expressed through
    tsb_bn_apply with scale 1 / shift 0 so it stays on the library path"""

    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        one = torch.ones(C, dtype=torch.float32, device=x.device)
        zero = torch.zeros(C, dtype=torch.float32, device=x.device)
        y = ops.nhwc_empty(N, C, H, W, device=x.device)
        ops.call("tsb_bn_apply", ops.ptr(x), ops.cs_of(x), ops.ptr(one), ops.ptr(zero), None, 0, 1, ops.ptr(y), C,
                 N * H * W, C, ops.stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return dy * (y > 0).to(dy.dtype)


class AttentionRefinement(nn.Module):
    """seg_oprs.py:192-212: fm = CBR3x3(x); fm * sigmoid(BN(conv1x1(GAP(fm)))).
    forward(x, add=None) additionally fuses the caller's `fm += last_fm` (bisenet network.py:91-92)."""

    def __init__(self, in_planes, out_planes, norm_layer=nn.BatchNorm2d):
        super(AttentionRefinement, self).__init__()
        self.conv_3x3 = ConvBnRelu(in_planes, out_planes, 3, 1, 1, has_bn=True, norm_layer=norm_layer,
                                   has_relu=True, has_bias=False)
        self.channel_attention = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            ConvBnRelu(out_planes, out_planes, 1, 1, 0, has_bn=True, norm_layer=norm_layer, has_relu=False,
                       has_bias=False),
            nn.Sigmoid())

    def forward(self, x, add=None):
        fm = self.conv_3x3(x)
        pooled = ops.AdaptiveAvgPoolFn.apply(fm, 1)
        a = self.channel_attention[1](pooled)  # pre-sigmoid logit; the sigmoid is fused into the scale kernel
        return ops.ChanScaleFn.apply(fm, a, None if add is None else _as_act(add), 0.0)


class FeatureFusion(nn.Module):
    """seg_oprs.py:215-238: fm = CBR1x1(cat[x1,x2]); fm + fm*sigmoid(conv1x1(relu(conv1x1(GAP(fm)))))"""

    def __init__(self, in_planes, out_planes, reduction=1, norm_layer=nn.BatchNorm2d):
        super(FeatureFusion, self).__init__()
        self.conv_1x1 = ConvBnRelu(in_planes, out_planes, 1, 1, 0, has_bn=True, norm_layer=norm_layer,
                                   has_relu=True, has_bias=False)
        self.channel_attention = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            ConvBnRelu(out_planes, out_planes // reduction, 1, 1, 0, has_bn=False, norm_layer=norm_layer,
                       has_relu=True, has_bias=False),
            ConvBnRelu(out_planes // reduction, out_planes, 1, 1, 0, has_bn=False, norm_layer=norm_layer,
                       has_relu=False, has_bias=False),
            nn.Sigmoid())

    def forward(self, x1, x2):
        fm = ops.ConcatFn.apply(_as_act(x1), _as_act(x2))
        fm = self.conv_1x1(fm)
        pooled = ops.AdaptiveAvgPoolFn.apply(fm, 1)
        a = self.channel_attention[2](self.channel_attention[1](pooled))
        return ops.ChanScaleFn.apply(fm, a, None, 1.0)


class GlobalAvgPool2d(nn.Module):
    """seg_oprs.py:97-107"""

    def forward(self, inputs):
        return ops.AdaptiveAvgPoolFn.apply(_as_act(inputs), 1)


def upsample_bilinear(x, size):
    """F.interpolate(x, size=size, mode='bilinear', align_corners=True) on the libtsb path"""
    return ops.BilinearFn.apply(_as_act(x), int(size[0]), int(size[1]))
