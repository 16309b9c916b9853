"""TEST INFRASTRUCTURE ONLY (oracle): CPU fp32 restatement of the reference network forward
`TextDetBase.forward` (basemodel.py:240-244) in plain torch.nn.functional calls, driven by
the same 3-key checkpoint dict.  It exists because /root/reference is not present on the
GPU box; `tests/test_oracle_vs_reference.py` pins it against the UNMODIFIED reference
(imported through oracle/ref_shim.py) in the build container, and tests/golden/ holds
vectors produced by the reference itself.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product (comic-text-detector_b200/) never does.

Reference map (file:line relative to /root/reference):
  * yolo graph walk                    models/yolov5/yolo.py:115-134 (+ cfg parse 208-259)
  * Conv = conv -> BN -> act           models/yolov5/common.py:30-49
  * BN folding for blk_det (eps 1e-3)  models/yolov5/yolo.py:185-192, utils/yolov5_utils.py:23-43,53-62
  * Bottleneck / C3 / SPPF / Concat    models/yolov5/common.py:94-104,126-138,181-196,267-274
  * Detect decode                      models/yolov5/yolo.py:23-44
  * UnetHead / double_conv(_up)_c3     basemodel.py:21-45,62-78
  * DBHead                             basemodel.py:106-125,130-157
"""
import torch
import torch.nn.functional as F

from oracle.synth import parse_cfg


def _act(x, kind):
    if kind == "silu":
        return F.silu(x)
    if kind == "leaky":
        return F.leaky_relu(x, 0.1)
    if kind == "relu":
        return F.relu(x)
    return x


class RefNet:
    """fp32 CPU forward. `calibrate_bn=True` re-estimates every BatchNorm's running stats
    from the activations of the given batch while running (used once, by
    oracle/synth.make_checkpoint, to make the random-weight net well conditioned)."""

    def __init__(self, ckpt):
        self.cfg = ckpt["blk_det"]["cfg"]
        self.layers = parse_cfg(self.cfg)
        self.yolo = ckpt["blk_det"]["weights"]
        self.seg = ckpt["text_seg"]
        self.det = ckpt["text_det"]
        self.calibrate_bn = False
        self._fused = {}

    # ---- building blocks ---------------------------------------------------------------
    def _bn_stats(self, sd, prefix, y):
        if self.calibrate_bn:
            sd[prefix + ".running_mean"] = y.mean((0, 2, 3)).detach().clone()
            sd[prefix + ".running_var"] = y.var((0, 2, 3), unbiased=False).detach().clone()

    def _yolo_conv(self, x, prefix, k, s, p=None):
        """Conv with BN folded at load like Model.fuse() (yolo.py:185-192): eps = 1e-3
        (yolov5_utils.py:59); fused weights = diag(g/sqrt(eps+var)) @ W (yolov5_utils.py:35-41)."""
        sd = self.yolo
        w = sd[prefix + ".conv.weight"]
        pad = k // 2 if p is None else p
        if self.calibrate_bn:
            self._bn_stats(sd, prefix + ".bn", F.conv2d(x, w, None, s, pad))
        if prefix not in self._fused or self.calibrate_bn:
            g, b = sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"]
            rm, rv = sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"]
            eps = 1e-3
            w_bn = torch.diag(g.div(torch.sqrt(eps + rv)))
            wf = torch.mm(w_bn, w.view(w.shape[0], -1)).view(w.shape)
            b_conv = torch.zeros(w.shape[0])
            b_bn = b - g.mul(rm).div(torch.sqrt(rv + eps))
            bf = torch.mm(w_bn, b_conv.reshape(-1, 1)).reshape(-1) + b_bn
            self._fused[prefix] = (wf, bf)
        wf, bf = self._fused[prefix]
        return F.silu(F.conv2d(x, wf, bf, s, pad))

    def _head_conv(self, sd, x, prefix, k, act):
        """heads' Conv: live BN eps 1e-5 (basemodel.py:223 fuse=False), act per `act` arg."""
        y = F.conv2d(x, sd[prefix + ".conv.weight"], None, 1, k // 2)
        self._bn_stats(sd, prefix + ".bn", y)
        y = F.batch_norm(y, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"],
                         sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], False, 0.0, 1e-5)
        return _act(y, act)

    def _c3(self, conv, x, prefix, n, shortcut):
        y1 = conv(x, prefix + ".cv1", 1)
        for j in range(n):
            t = conv(conv(y1, "%s.m.%d.cv1" % (prefix, j), 1), "%s.m.%d.cv2" % (prefix, j), 3)
            y1 = y1 + t if shortcut else t
        y2 = conv(x, prefix + ".cv2", 1)
        return conv(torch.cat((y1, y2), 1), prefix + ".cv3", 1)

    # ---- yolo -----------------------------------------------------------------------------
    def yolo_forward(self, x):
        conv = lambda t, pfx, k, s=1, p=None: self._yolo_conv(t, pfx, k, s, p)
        outs = []
        feats = []
        for L in self.layers:
            i, f, t = L["i"], L["f"], L["type"]
            if f != -1:
                x = outs[f] if isinstance(f, int) else [x if j == -1 else outs[j] for j in f]
            pfx = "model.%d" % i
            if t == "Conv":
                x = conv(x, pfx, L["k"], L["s"], L["p"])
            elif t == "C3":
                x = self._c3(lambda a, b, k: conv(a, b, k), x, pfx, L["n"], L["shortcut"])
            elif t == "SPPF":
                x = conv(x, pfx + ".cv1", 1)
                y1 = F.max_pool2d(x, L["k"], 1, L["k"] // 2)
                y2 = F.max_pool2d(y1, L["k"], 1, L["k"] // 2)
                y3 = F.max_pool2d(y2, L["k"], 1, L["k"] // 2)
                x = conv(torch.cat([x, y1, y2, y3], 1), pfx + ".cv2", 1)
            elif t == "Upsample":
                x = F.interpolate(x, scale_factor=2, mode="nearest")
            elif t == "Concat":
                x = torch.cat(x, 1)
            elif t == "Detect":
                x = self._detect(x, pfx, L)
            outs.append(x)
            if i in (1, 3, 5, 7, 9):
                feats.append(x)
        return x, feats

    def _detect(self, xs, pfx, L):
        sd = self.yolo
        nc = L["nc"]
        no, na = nc + 5, len(L["anchors"][0]) // 2
        anchors = sd[pfx + ".anchors"]  # (nl, na, 2), already / stride
        z = []
        for li, x in enumerate(xs):
            stride = float(8 * 2 ** li)
            y = F.conv2d(x, sd["%s.m.%d.weight" % (pfx, li)], sd["%s.m.%d.bias" % (pfx, li)])
            bs, _, ny, nx = y.shape
            y = y.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
            grid = torch.stack((xv, yv), 2).expand((1, na, ny, nx, 2)).float()
            anchor_grid = (anchors[li].clone() * stride).view((1, na, 1, 1, 2)).expand((1, na, ny, nx, 2)).float()
            y = y.sigmoid()
            y[..., 0:2] = (y[..., 0:2] * 2 - 0.5 + grid) * stride
            y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anchor_grid
            z.append(y.view(bs, -1, no))
        return torch.cat(z, 1)

    # ---- heads ------------------------------------------------------------------------------
    def _up_c3(self, sd, x, pfx):
        """double_conv_up_c3 (basemodel.py:21-32): C3(leaky) -> ConvT4x4s2p1 -> BN -> ReLU."""
        conv = lambda a, b, k: self._head_conv(sd, a, b, k, "leaky")
        x = self._c3(conv, x, pfx + ".conv.0", 1, True)
        y = F.conv_transpose2d(x, sd[pfx + ".conv.1.weight"], None, 2, 1)
        self._bn_stats(sd, pfx + ".conv.2", y)
        y = F.batch_norm(y, sd[pfx + ".conv.2.running_mean"], sd[pfx + ".conv.2.running_var"],
                         sd[pfx + ".conv.2.weight"], sd[pfx + ".conv.2.bias"], False, 0.0, 1e-5)
        return F.relu(y)

    def seg_forward(self, f256, f128, f64, f32, f3):
        sd = self.seg
        conv = lambda a, b, k: self._head_conv(sd, a, b, k, "leaky")
        d16 = self._c3(conv, F.avg_pool2d(f3, 2, 2), "down_conv1.conv", 1, True)
        u32 = self._up_c3(sd, d16, "upconv0")
        u64 = self._up_c3(sd, torch.cat([f32, u32], 1), "upconv2")
        u128 = self._up_c3(sd, torch.cat([f64, u64], 1), "upconv3")
        u256 = self._up_c3(sd, torch.cat([f128, u128], 1), "upconv4")
        u512 = self._up_c3(sd, torch.cat([f256, u256], 1), "upconv5")
        logit = F.conv_transpose2d(u512, sd["upconv6.0.weight"], None, 2, 1)
        self.last_seg_logit = logit
        return torch.sigmoid(logit), [f128, f64, u64]

    def _db_tail(self, sd, x, name):
        b0 = sd.get(name + ".0.bias")
        y = F.conv2d(x, sd[name + ".0.weight"], b0, 1, 1)
        self._bn_stats(sd, name + ".1", y)
        y = F.relu(F.batch_norm(y, sd[name + ".1.running_mean"], sd[name + ".1.running_var"],
                                sd[name + ".1.weight"], sd[name + ".1.bias"], False, 0.0, 1e-5))
        y = F.conv_transpose2d(y, sd[name + ".3.weight"], sd[name + ".3.bias"], 2)
        self._bn_stats(sd, name + ".4", y)
        y = F.relu(F.batch_norm(y, sd[name + ".4.running_mean"], sd[name + ".4.running_var"],
                                sd[name + ".4.weight"], sd[name + ".4.bias"], False, 0.0, 1e-5))
        return F.conv_transpose2d(y, sd[name + ".6.weight"], sd[name + ".6.bias"], 2)

    def det_forward(self, f128, f64, u64):
        sd = self.det
        u128 = self._up_c3(sd, torch.cat([f64, u64], 1), "upconv3")
        x = self._up_c3(sd, torch.cat([f128, u128], 1), "upconv4")
        y = F.conv2d(x, sd["conv.0.weight"], sd["conv.0.bias"])
        self._bn_stats(sd, "conv.1", y)
        x = F.relu(F.batch_norm(y, sd["conv.1.running_mean"], sd["conv.1.running_var"],
                                sd["conv.1.weight"], sd["conv.1.bias"], False, 0.0, 1e-5))
        thr_logit = self._db_tail(sd, x, "thresh")
        bin_logit = self._db_tail(sd, x, "binarize")
        self.last_thr_logit, self.last_bin_logit = thr_logit, bin_logit
        return torch.cat((torch.sigmoid(bin_logit), torch.sigmoid(thr_logit)), 1)

    @torch.no_grad()
    def forward(self, x, calibrate_bn=False):
        """x: f32 (N,3,H,W) BGR in [0,1] -> (blks (N,A,7), mask (N,1,H,W), lines (N,2,H,W))."""
        self.calibrate_bn = calibrate_bn
        blks, feats = self.yolo_forward(x)
        self.last_feats = feats
        mask, feats2 = self.seg_forward(*feats)
        lines = self.det_forward(*feats2)
        self.calibrate_bn = False
        return blks, mask, lines

    __call__ = forward
