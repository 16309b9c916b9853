"""TEST INFRASTRUCTURE: one-off search for the frozen constants in oracle/synth.py:CALIB.
Needs /root/reference (runs the unmodified reference forward).  Usage:
    python -m oracle.calibrate        # prints the dict to paste into synth.py
"""
import sys
import tempfile

import numpy as np
import torch

from oracle import ref_shim, synth


def logit(p):
    return float(np.log(p / (1 - p)))


SMOOTH = "--rough" not in sys.argv


def main():
    ns = ref_shim.load()
    ck = synth.make_checkpoint(0, smooth=SMOOTH, calib=synth.CALIB_IDENTITY)
    f = tempfile.NamedTemporaryFile(suffix=".ckpt", delete=False).name
    torch.save(ck, f)
    net = ns.basemodel.TextDetBase(f, device="cpu", act="leaky")
    img = synth.structured_page(1000)
    x = torch.from_numpy(img.transpose(2, 0, 1)[None].astype(np.float32) / 255)
    cap = {}

    def hook(name):
        def fn(m, i, o):
            cap[name] = o.detach()
        return fn
    net.text_seg.upconv6[0].register_forward_hook(hook("seg"))
    net.text_det.binarize[6].register_forward_hook(hook("bin"))
    net.text_det.thresh[6].register_forward_hook(hook("thr"))
    for i, m in enumerate(net.blk_det.model[24].m):
        m.register_forward_hook(hook("det%d" % i))
    with torch.no_grad():
        net(x)
    out = {}
    seg = cap["seg"].numpy().ravel()
    out["seg_gain"] = round(2.5 / float(seg.std()), 3)
    out["seg_bias"] = round(-float(np.quantile(seg * out["seg_gain"], 0.85)), 3)  # 15 % above 0.5
    b = cap["bin"].numpy().ravel() - float(net.text_det.binarize[6].bias)
    out["db_bin_gain"] = round(2.5 / float(b.std()), 3)
    out["db_bin_bias"] = round(logit(0.3) - float(np.quantile(b * out["db_bin_gain"], 0.88)), 3)
    t = cap["thr"].numpy().ravel() - float(net.text_det.thresh[6].bias)
    out["db_thr_gain"] = round(1.0 / float(t.std()), 3)
    out["db_thr_bias"] = round(-float(t.mean()) * out["db_thr_gain"], 3)
    objs, obj_bias = [], []
    for i in range(3):
        d = cap["det%d" % i][0].view(3, 7, -1)
        bias = net.blk_det.model[24].m[i].bias.detach().view(3, 7)
        objs.append((d[:, 4] - bias[:, 4:5]).detach().numpy().ravel())
    gain = 2.0 / float(np.concatenate(objs).std())
    out["det_obj_gain"] = round(gain, 3)
    for i, frac in enumerate((0.0015, 0.004, 0.012)):  # share of anchors with obj > 0.5 per level
        obj_bias.append(round(-float(np.quantile(objs[i] * gain, 1 - frac)), 3))
    out["det_obj_bias"] = obj_bias
    d = cap["det0"][0].view(3, 7, -1)
    bias = net.blk_det.model[24].m[0].bias.detach().view(3, 7)
    out["det_cls_gain"] = round(1.5 / float((d[:, 5:] - bias[:, 5:, None]).std()), 3)
    out["det_cls_bias"] = 2.0
    out["det_box_gain"] = round(0.6 / float((d[:, :4] - bias[:, :4, None]).std()), 3)
    print("CALIB = dict(CALIB_IDENTITY)\nCALIB.update(%r)" % (out,))
    return out


if __name__ == "__main__":
    main()
