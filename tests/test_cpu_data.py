"""CPU tests of the input-pipeline oracle (oracle/data_ref.py) and the host side of the device pipeline
(torchseg_b200/utils/gpu_pipeline.py): OpenCV's fixed-point resize restated bit for bit, the whole TrainPre sequence
against fixtures generated from the LIVE reference (tests/golden/data_pipeline.json), and the random-draw order."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from golden_cases import pipeline_case, pipeline_case_dfn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "data_pipeline.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_resize_restatements_match_cv2_bit_for_bit():
    cv2 = pytest.importorskip("cv2")
    from oracle import data_ref
    rng = np.random.default_rng(0)
    for (H, W) in ((64, 128), (101, 77), (250, 333)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        gt = rng.integers(0, 20, (H, W), dtype=np.uint8)
        for scale in (0.5, 0.75, 1, 1.25, 1.5, 1.75, 2.0):          # bisenet config.py:86 / dfn config.py:87
            sh, sw = int(H * scale), int(W * scale)
            for ipp in (True, False):
                if hasattr(cv2, "ipp"):
                    cv2.ipp.setUseIPP(ipp)
                assert np.array_equal(cv2.resize(img, (sw, sh), interpolation=cv2.INTER_LINEAR),
                                      data_ref.resize_linear_u8(img, sw, sh)), (H, W, scale, ipp)
            assert np.array_equal(cv2.resize(gt, (sw, sh), interpolation=cv2.INTER_NEAREST),
                                  data_ref.resize_nearest(gt, sw, sh)), (H, W, scale)
    if hasattr(cv2, "ipp"):
        cv2.ipp.setUseIPP(True)


def test_train_pre_oracle_vs_live_reference_golden():
    from oracle import data_ref
    bgr, gt, crop, scales, mean, std = pipeline_case()
    seen = set()
    for seed, ent in GOLD["cases"].items():
        random.seed(int(seed))
        prm = data_ref.draw_params(bgr.shape[:2], crop, scales)
        data, label = data_ref.train_pre(bgr[:, :, ::-1], gt, prm, crop, mean, std)
        assert list(data.shape) == ent["shape"] and data.dtype == np.float32 and label.dtype == np.int64
        assert _sha(data) == ent["data_sha256"], seed
        assert _sha(label) == ent["label_sha256"], seed
        seen.add((prm["flip"], prm["scale"]))
    assert len(seen) >= 8          # the seeds cover both mirror states and all six scales incl. the padded 0.75 case


def test_host_draw_order_and_lut_match_reference_semantics():
    from oracle import data_ref
    from torchseg_b200.utils.gpu_pipeline import TrainPreGPU, normalize_lut
    bgr, gt, crop, scales, mean, std = pipeline_case()
    pre = TrainPreGPU.__new__(TrainPreGPU)                         # host logic only: no device needed
    pre.crop_h, pre.crop_w, pre.scale_array = crop[0], crop[1], list(scales)
    for seed in range(32):
        random.seed(seed)
        a = pre.draw(bgr.shape[:2])
        state_a = random.getstate()
        random.seed(seed)
        b = data_ref.draw_params(bgr.shape[:2], crop, scales)
        assert random.getstate() == state_a                        # same number of draws consumed
        assert (a["flip"], a["sh"], a["sw"], a["pos_h"], a["pos_w"]) == (b["flip"], b["sh"], b["sw"], b["pos_h"], b["pos_w"])
    # normalize(): every uint8 value through the reference's float32 / float64 sequence (img_utils.py:181-187)
    lut = normalize_lut(mean, std)
    v = np.arange(256, dtype=np.uint8).reshape(256, 1, 1).repeat(3, axis=2)
    ref = (((v.astype(np.float32) / 255.0) - mean) / std).astype(np.float32)      # [256,1,3]
    assert np.array_equal(lut.view(np.uint32), np.ascontiguousarray(ref[:, 0, :].T).view(np.uint32))


def test_canny_dilate_restatements_match_cv2_bit_for_bit():
    """cv2.Canny(apertureSize=7) / cv2.dilate(7x7) as used by the DFN loader (dfn dataloader.py:22-29)"""
    cv2 = pytest.importorskip("cv2")
    from oracle import data_ref
    rng = np.random.default_rng(3)
    k = cv2.getStructuringElement(cv2.MORPH_RECT, (7, 7))
    for (H, W) in ((64, 96), (101, 77), (200, 312)):
        for _ in range(2):
            g = np.zeros((H, W), np.uint8)
            for _ in range(40):
                y0, x0 = rng.integers(0, H), rng.integers(0, W)
                g[y0:y0 + rng.integers(3, H // 2), x0:x0 + rng.integers(3, W // 2)] = rng.integers(0, 19)
            g[rng.random((H, W)) < 0.01] = rng.integers(0, 19)
            ref = cv2.Canny(g, 5, 5, apertureSize=7)
            mine = data_ref.canny_ap7(g, 5, 5)
            assert np.array_equal(ref, mine) and int((ref > 0).sum()) > 100
            assert np.array_equal(cv2.dilate(ref, k), data_ref.dilate7(mine))
            dx = cv2.Sobel(g, cv2.CV_16S, 1, 0, ksize=7, scale=1 / 16.0, borderType=cv2.BORDER_REPLICATE)
            assert np.array_equal(dx.astype(np.int64), data_ref.sobel7_over16(g)[0])


def test_train_pre_dfn_oracle_vs_live_reference_golden():
    from oracle import data_ref
    bgr, gt, crop, scales, mean, std = pipeline_case_dfn()
    for seed, ent in GOLD["cases_dfn"].items():
        random.seed(int(seed))
        prm = data_ref.draw_params(bgr.shape[:2], crop, scales)
        data, label, aux = data_ref.train_pre_dfn(bgr[:, :, ::-1], gt, prm, crop, mean, std)
        assert _sha(data) == ent["data_sha256"] and _sha(label) == ent["label_sha256"], seed
        assert _sha(aux) == ent["aux_sha256"] and int((aux == 1).sum()) == ent["aux_ones"], seed


def test_train_pre_speed_config_label_downsampling_vs_live_reference_golden():
    from oracle import data_ref
    bgr, gt, crop, scales, mean, std = pipeline_case()
    for seed, ent in GOLD["cases_speed"].items():
        random.seed(int(seed))
        prm = data_ref.draw_params(bgr.shape[:2], crop, scales)
        _, label = data_ref.train_pre(bgr[:, :, ::-1], gt, prm, crop, mean, std, gt_down_sampling=8)
        assert list(label.shape) == ent["label_shape"] == [crop[0] // 8, crop[1] // 8]
        assert _sha(label) == ent["label_sha256"], seed
