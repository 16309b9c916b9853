"""TEST INFRASTRUCTURE (oracle side): synthetic checkpoint + synthetic pages.

The reference ships no weights (README.md:5 points at external downloads) and its
YOLOv5s cfg travels inside the checkpoint (models/yolov5/yolo.py:292), so every parity /
bench run uses a seeded random-weight checkpoint in the reference's own 3-key format
(utils/export.py:23-28; loaded by basemodel.py:211-217):

    {'blk_det': {'cfg': <yolov5s v6.0 dict, nc=2>, 'weights': state_dict},
     'text_seg': state_dict(UnetHead), 'text_det': state_dict(DBHead)}

Plain default init gives degenerate outputs (flat 0.5 maps, 0 boxes), so
after seeding we (a) randomise every BatchNorm's affine + running stats so BN folding is
really exercised, and (b) rescale/bias the last layers with FROZEN constants (found once
with oracle/calibrate.py and pasted below) so that the DB map has blobs, the seg mask
crosses 0.5 and some boxes survive conf 0.4.

This file only uses torch (no reference import), so it also runs on the GPU box.
"""
import copy
import math

import numpy as np
import torch
import torch.nn as nn

YOLOV5S_CFG = {
    "nc": 2, "depth_multiple": 0.33, "width_multiple": 0.50,
    "anchors": [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
    "backbone": [[-1, 1, "Conv", [64, 6, 2, 2]], [-1, 1, "Conv", [128, 3, 2]], [-1, 3, "C3", [128]],
                 [-1, 1, "Conv", [256, 3, 2]], [-1, 6, "C3", [256]], [-1, 1, "Conv", [512, 3, 2]],
                 [-1, 9, "C3", [512]], [-1, 1, "Conv", [1024, 3, 2]], [-1, 3, "C3", [1024]],
                 [-1, 1, "SPPF", [1024, 5]]],
    "head": [[-1, 1, "Conv", [512, 1, 1]], [-1, 1, "nn.Upsample", [None, 2, "nearest"]],
             [[-1, 6], 1, "Concat", [1]], [-1, 3, "C3", [512, False]],
             [-1, 1, "Conv", [256, 1, 1]], [-1, 1, "nn.Upsample", [None, 2, "nearest"]],
             [[-1, 4], 1, "Concat", [1]], [-1, 3, "C3", [256, False]],
             [-1, 1, "Conv", [256, 3, 2]], [[-1, 14], 1, "Concat", [1]], [-1, 3, "C3", [512, False]],
             [-1, 1, "Conv", [512, 3, 2]], [[-1, 10], 1, "Concat", [1]], [-1, 3, "C3", [1024, False]],
             [[17, 20, 23], 1, "Detect", ["nc", "anchors"]]],
}

# ---- frozen calibration constants (oracle/calibrate.py prints these) -------------------
CALIB_IDENTITY = {
    "seg_gain": 1.0, "seg_bias": 0.0,
    "db_bin_gain": 1.0, "db_bin_bias": 0.0,
    "db_thr_gain": 1.0, "db_thr_bias": 0.0,
    "det_obj_gain": 1.0, "det_obj_bias": [0.0, 0.0, 0.0],
    "det_cls_gain": 1.0, "det_cls_bias": 0.0, "det_box_gain": 1.0,
}
CALIB = dict(CALIB_IDENTITY)
CALIB.update({'seg_gain': 4.188, 'seg_bias': -5.886, 'db_bin_gain': 4.763, 'db_bin_bias': -6.833,
              'db_thr_gain': 1.803, 'db_thr_bias': -1.036, 'det_obj_gain': 5.273,
              'det_obj_bias': [-12.458, -9.706, -6.742], 'det_cls_gain': 3.714, 'det_cls_bias': 2.0,
              'det_box_gain': 0.5})


# ---- minimal module mirrors used ONLY to obtain state_dicts with the reference's key
# layout and torch's default initialisers; arithmetic lives elsewhere.
def _make_divisible(x, d):
    return math.ceil(x / d) * d


class _Conv(nn.Module):
    def __init__(self, c1, c2, k=1, s=1, p=None):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2 if p is None else p, bias=False)
        self.bn = nn.BatchNorm2d(c2)


class _Bottleneck(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1 = _Conv(c1, c2, 1, 1)
        self.cv2 = _Conv(c2, c2, 3, 1)


class _C3(nn.Module):
    def __init__(self, c1, c2, n=1):
        super().__init__()
        c_ = int(c2 * 0.5)
        self.cv1 = _Conv(c1, c_, 1, 1)
        self.cv2 = _Conv(c1, c_, 1, 1)
        self.cv3 = _Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(_Bottleneck(c_, c_) for _ in range(n)))


class _SPPF(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1 = _Conv(c1, c1 // 2, 1, 1)
        self.cv2 = _Conv(c1 // 2 * 4, c2, 1, 1)


class _Detect(nn.Module):
    def __init__(self, nc, anchors, ch, strides):
        super().__init__()
        na = len(anchors[0]) // 2
        a = torch.tensor(anchors).float().view(len(anchors), -1, 2)
        self.register_buffer("anchors", a / torch.tensor(strides).float().view(-1, 1, 1))
        self.m = nn.ModuleList(nn.Conv2d(x, (nc + 5) * na, 1) for x in ch)


class _Holder(nn.Module):
    pass


def parse_cfg(cfg):
    """Channel/topology walk of the yolov5 cfg (models/yolov5/yolo.py:208-259 semantics for
    the 6 module kinds the shipped cfg uses).  Returns a list of dicts."""
    gd, gw = cfg["depth_multiple"], cfg["width_multiple"]
    nc, anchors = cfg["nc"], cfg["anchors"]
    na = len(anchors[0]) // 2
    no = na * (nc + 5)
    ch = [3]
    layers = []
    for i, (f, n, m, args) in enumerate(cfg["backbone"] + cfg["head"]):
        args = [nc if a == "nc" else anchors if a == "anchors" else a for a in args]
        n = max(round(n * gd), 1) if n > 1 else n
        if m in ("Conv", "C3", "SPPF"):
            c1, c2 = ch[f], args[0]
            if c2 != no:
                c2 = _make_divisible(c2 * gw, 8)
            L = dict(i=i, f=f, type=m, c1=c1, c2=c2)
            if m == "Conv":
                L.update(k=args[1] if len(args) > 1 else 1, s=args[2] if len(args) > 2 else 1,
                         p=args[3] if len(args) > 3 else None)
            elif m == "C3":
                L.update(n=n, shortcut=args[1] if len(args) > 1 else True)
            elif m == "SPPF":
                L.update(k=args[1])
        elif m == "Concat":
            c2 = sum(ch[x] for x in f)
            L = dict(i=i, f=f, type=m, c2=c2)
        elif m == "Detect":
            L = dict(i=i, f=f, type=m, nc=nc, anchors=anchors, ch=[ch[x] for x in f], c2=None)
            c2 = None
        elif m == "nn.Upsample":
            c2 = ch[f]
            L = dict(i=i, f=f, type="Upsample", c2=c2)
        else:
            raise ValueError("cfg module %r is outside the shipped yolov5s cfg" % m)
        layers.append(L)
        if i == 0:
            ch = []
        ch.append(c2)
    return layers


def _build_yolo_holder(cfg):
    layers = parse_cfg(cfg)
    model = nn.ModuleList()
    for L in layers:
        t = L["type"]
        if t == "Conv":
            model.append(_Conv(L["c1"], L["c2"], L["k"], L["s"], L["p"]))
        elif t == "C3":
            model.append(_C3(L["c1"], L["c2"], L["n"]))
        elif t == "SPPF":
            model.append(_SPPF(L["c1"], L["c2"]))
        elif t == "Detect":
            model.append(_Detect(L["nc"], L["anchors"], L["ch"], [8., 16., 32.]))
        else:
            model.append(nn.Identity())
    h = _Holder()
    h.model = model
    return h


class _UpC3(nn.Module):  # double_conv_up_c3 (basemodel.py:21-32)
    def __init__(self, in_ch, mid_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(_C3(in_ch + mid_ch, mid_ch), nn.ConvTranspose2d(mid_ch, out_ch, 4, 2, 1, bias=False),
                                  nn.BatchNorm2d(out_ch), nn.ReLU())


class _DownC3(nn.Module):  # double_conv_c3 (basemodel.py:34-45)
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = _C3(in_ch, out_ch)


def _build_unet_holder():  # UnetHead (basemodel.py:47-60)
    h = _Holder()
    h.down_conv1 = _DownC3(512, 512)
    h.upconv0 = _UpC3(0, 512, 256)
    h.upconv2 = _UpC3(256, 512, 256)
    h.upconv3 = _UpC3(0, 512, 256)
    h.upconv4 = _UpC3(128, 256, 128)
    h.upconv5 = _UpC3(64, 128, 64)
    h.upconv6 = nn.Sequential(nn.ConvTranspose2d(64, 1, 4, 2, 1, bias=False), nn.Sigmoid())
    return h


def _build_db_holder(c=64):  # DBHead (basemodel.py:83-104,130-157)
    h = _Holder()
    h.upconv3 = _UpC3(0, 512, 256)
    h.upconv4 = _UpC3(128, 256, 128)
    h.conv = nn.Sequential(nn.Conv2d(128, c, 1), nn.BatchNorm2d(c), nn.ReLU())
    h.binarize = nn.Sequential(nn.Conv2d(c, c // 4, 3, padding=1), nn.BatchNorm2d(c // 4), nn.ReLU(),
                               nn.ConvTranspose2d(c // 4, c // 4, 2, 2), nn.BatchNorm2d(c // 4), nn.ReLU(),
                               nn.ConvTranspose2d(c // 4, 1, 2, 2))
    h.thresh = nn.Sequential(nn.Conv2d(c, c // 4, 3, padding=1, bias=False), nn.BatchNorm2d(c // 4), nn.ReLU(),
                             nn.ConvTranspose2d(c // 4, c // 4, 2, 2), nn.BatchNorm2d(c // 4), nn.ReLU(),
                             nn.ConvTranspose2d(c // 4, 1, 2, 2), nn.Sigmoid())
    return h


def _randomise_bn(mod, g):
    for m in mod.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            m.weight.data = torch.empty(n).uniform_(0.5, 1.5, generator=g)
            m.bias.data = torch.randn(n, generator=g) * 0.3
            m.running_mean.data = torch.randn(n, generator=g) * 0.1
            m.running_var.data = torch.empty(n).uniform_(0.5, 1.5, generator=g)


_BILINEAR = torch.tensor([1.0, 3.0, 3.0, 1.0]) / 4.0


def _smooth_deconvs(mod, g):
    """Random-weight ConvTranspose stacks emit a 2x2/4x4 checkerboard (every pixel its own
    contour: 130k contours/page), which exercises none of the box logic.
    For the 'smooth' checkpoint every ConvT4x4s2 becomes (random channel mix) x (bilinear
    x2 kernel) and every ConvT2x2s2 (random channel mix) x (constant 2x2) so the maps are
    blob-structured.  The 'rough' checkpoint keeps fully random taps (used for logits parity,
    where a swapped tap must show up)."""
    for m in mod.modules():
        if isinstance(m, nn.ConvTranspose2d):
            w = m.weight.data
            k = w.shape[-1]
            a = torch.randn(w.shape[0], w.shape[1], 1, 1, generator=g) / math.sqrt(w.shape[0])
            if k == 4:
                m.weight.data = (a * (_BILINEAR[:, None] * _BILINEAR[None, :])).contiguous()
            else:
                m.weight.data = a.expand(-1, -1, k, k).contiguous()


def make_checkpoint(seed: int = 0, smooth: bool = True, calib=None, bn_calibrate: int = 1024):
    """Seeded synthetic checkpoint dict in the reference's 3-key format."""
    if calib is None:
        calib = CALIB if smooth else CALIB_IDENTITY
    torch.manual_seed(seed)
    yolo = _build_yolo_holder(YOLOV5S_CFG)
    seg = _build_unet_holder()
    db = _build_db_holder(64)
    g = torch.Generator().manual_seed(seed + 12345)
    for mod in (yolo, seg, db):
        _randomise_bn(mod, g)
    if smooth:
        _smooth_deconvs(seg, g)
        _smooth_deconvs(db, g)
    ck = {
        "blk_det": {"cfg": copy.deepcopy(YOLOV5S_CFG), "weights": yolo.state_dict()},
        "text_seg": seg.state_dict(), "text_det": db.state_dict(),
    }
    if bn_calibrate:
        # Re-estimate every BN's running stats on one synthetic page, as training would have:
        # without it the random net is ill conditioned (tiny spatial signal on a large DC; the
        # last-layer gains come out ~500 and amplify rounding noise far above any tolerance).
        from oracle.net_ref import RefNet  # lazy: net_ref imports this module
        page = structured_page(1000 + seed, bn_calibrate, bn_calibrate)
        x = torch.from_numpy(page.transpose(2, 0, 1)[None].astype(np.float32) / 255)
        RefNet(ck)(x, calibrate_bn=True)  # writes running_mean/var into the three state dicts
        yolo.load_state_dict(ck["blk_det"]["weights"])
        seg.load_state_dict(ck["text_seg"])
        db.load_state_dict(ck["text_det"])
    # --- calibration of the last layers ---------------------------------------------
    with torch.no_grad():
        # seg: upconv6 has no bias (basemodel.py:58); channel 0 of u512 is turned into the
        # constant 1 (BN gamma=0, beta=1 -> ReLU -> 1) and its deconv taps carry the bias.
        seg.upconv6[0].weight.mul_(calib["seg_gain"])
        if calib["seg_bias"] != 0.0:
            bn = seg.upconv5.conv[2]
            bn.weight[0] = 0.0
            bn.bias[0] = 1.0
            seg.upconv6[0].weight[0, 0] = calib["seg_bias"] * (_BILINEAR[:, None] * _BILINEAR[None, :])
        db.binarize[6].weight.mul_(calib["db_bin_gain"])
        db.binarize[6].bias.fill_(calib["db_bin_bias"])
        db.thresh[6].weight.mul_(calib["db_thr_gain"])
        db.thresh[6].bias.fill_(calib["db_thr_bias"])
        det = yolo.model[24]
        for li, mi in enumerate(det.m):
            w = mi.weight.view(3, 7, -1)
            b = mi.bias.view(3, 7)
            w[:, :4] *= calib["det_box_gain"]
            w[:, 4] *= calib["det_obj_gain"]
            w[:, 5:] *= calib["det_cls_gain"]
            b[:, :2] = 0.0
            b[:, 2:4] = torch.randn(3, 2, generator=g) * 0.4  # per-anchor size variety
            b[:, 4] = calib["det_obj_bias"][li]
            b[:, 5:] = calib["det_cls_bias"]
    return {
        "blk_det": {"cfg": copy.deepcopy(YOLOV5S_CFG), "weights": {k: v.clone() for k, v in yolo.state_dict().items()}},
        "text_seg": {k: v.clone() for k, v in seg.state_dict().items()},
        "text_det": {k: v.clone() for k, v in db.state_dict().items()},
    }


# ---------------------------------------------------------------------------------------
# synthetic pages


def noise_page(seed: int, h: int = 1024, w: int = 1024) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def structured_page(seed: int, h: int = 1024, w: int = 1024) -> np.ndarray:
    """white page + random filled/outlined rectangles ('panels', 'bubbles') + glyph strings."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.full((h, w, 3), 255, np.uint8)
    for _ in range(60):
        x0, y0 = int(rng.integers(0, w - 40)), int(rng.integers(0, h - 40))
        x1, y1 = min(w - 1, x0 + int(rng.integers(20, 300))), min(h - 1, y0 + int(rng.integers(20, 300)))
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        if rng.random() < 0.5:
            cv2.rectangle(img, (x0, y0), (x1, y1), col, -1)
        else:
            cv2.rectangle(img, (x0, y0), (x1, y1), col, int(rng.integers(1, 6)))
    chars = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789!?"
    for _ in range(120):
        n = int(rng.integers(2, 12))
        s = "".join(chars[int(i)] for i in rng.integers(0, len(chars), n))
        org = (int(rng.integers(0, w - 100)), int(rng.integers(20, h - 10)))
        col = tuple(int(c) for c in rng.integers(0, 120, 3))
        cv2.putText(img, s, org, cv2.FONT_HERSHEY_SIMPLEX, float(rng.uniform(0.4, 1.6)), col,
                    int(rng.integers(1, 4)), cv2.LINE_AA)
    return img
