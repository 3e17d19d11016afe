"""Seeded synthetic cases shared by tools/make_golden.py (live reference) and the parity tests."""
import torch
import torch.nn.functional as F

OHEM_REGIMES = ["A", "B", "C", "D", "E"]


def ohem_case(regime, N=2, C=19, H=48, W=64, seed=1234):
    """SURVEY.md §8d regimes: A all kept (thr 0.7), B k-th value threshold, C min_kept > valid, D min_kept = 0,
    E all ignored"""
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, C, (N, H, W), generator=g, dtype=torch.int64)
    labels[:, : H // 10 + 1, :] = 255
    logits = torch.randn(N, C, H, W, generator=g)
    min_kept = N * H * W // 16
    if regime == "B":
        onehot = F.one_hot(labels.clamp(max=C - 1), C).permute(0, 3, 1, 2).float()
        logits = 8 * onehot + logits
    elif regime == "C":
        min_kept = N * H * W
    elif regime == "D":
        min_kept = 0
    elif regime == "E":
        labels[:] = 255
    return logits, labels, min_kept


def bisenet_case(N=2, HW=128, seed=77):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(0, 19, (N, HW, HW), generator=g, dtype=torch.int64)
    y[:, : HW // 10, :] = 255
    return x, y, N * HW * HW // 16, seed


def fcn_case(N=2, HW=256, seed=304):
    """BASELINE configs[0]: 2 synthetic 256x256 19-class images"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(0, 19, (N, HW, HW), generator=g, dtype=torch.int64)
    y[:, : HW // 10, :] = 255
    return x, y, seed


def fcn_r101_case(N=2, HW=128, seed=304, classes=21):
    """the shipped FCN-32s (R101_v1c, VOC 21 classes, ignore 255) at a small size"""
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(0, classes, (N, HW, HW), generator=g, dtype=torch.int64)
    y[:, : HW // 10, :] = 255
    return x, y, seed


def pspnet_case(N=2, HW=96, seed=4, classes=150):
    """BASELINE configs[2] shape family (PSPNet R101_v1c dilated-8, ADE 150 classes, ignore -1), small spatial size"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(-1, classes, (N, HW, HW), generator=g, dtype=torch.int64)
    return x, y, seed


def dfn_case(N=2, HW=128, seed=12345):
    """BASELINE configs C4 shape family (DFN R101_v1c, Cityscapes 19 classes + {0,1,255} border labels), small size"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(0, 19, (N, HW, HW), generator=g, dtype=torch.int64)
    y[:, : HW // 10, :] = 255
    e = (torch.rand(N, HW, HW, generator=g) < 0.15).to(torch.int64)
    e[:, : HW // 10, :] = 255
    return x, y, e, seed


def psanet_case(N=1, HW=480, seed=304, classes=150):
    """SURVEY C5 (PSANet R101_v1c dilated-8, ADE 150 classes, ignore -1). The PSA head needs h*w = 3600 positions at
    1/8 resolution, so the spatial size cannot be reduced below the reference's 480x480 (psanet config.py)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 3, HW, HW, generator=g)
    y = torch.randint(-1, classes, (N, HW, HW), generator=g, dtype=torch.int64)
    return x, y, seed


def pipeline_case(H=100, W=160, crop=(96, 128), seed=1):
    """a decoded uint8 BGR frame + label map, the BiSeNet scale array and normalisation constants (config.py:61-62,86)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    gt = rng.integers(0, 20, (H, W), dtype=np.uint8)
    gt[rng.random((H, W)) < 0.05] = 255
    return (bgr, gt, crop, [0.75, 1, 1.25, 1.5, 1.75, 2.0], np.array([0.485, 0.456, 0.406]),
            np.array([0.229, 0.224, 0.225]))


def pipeline_case_dfn(H=100, W=160, crop=(96, 128), seed=2):
    """label map with regions (so that Canny finds borders) + the DFN scale array (dfn config.py:87)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    gt = np.zeros((H, W), np.uint8)
    for _ in range(30):
        y0, x0 = rng.integers(0, H), rng.integers(0, W)
        h, w = rng.integers(3, H // 2), rng.integers(3, W // 2)
        gt[y0:y0 + h, x0:x0 + w] = rng.integers(0, 19)
    gt[rng.random((H, W)) < 0.03] = 255
    return (bgr, gt, crop, [0.5, 0.75, 1, 1.5, 1.75, 2.0], np.array([0.485, 0.456, 0.406]),
            np.array([0.229, 0.224, 0.225]))
