"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's post-processing stages
that the GPU kernels replace, each citing the reference lines it follows.  cv2 / torchvision are
the same library builds on the GPU box (same image), so they are used directly where the
reference itself calls them.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs import this.
"""
import numpy as np
import torch
import torchvision


def xywh2xyxy(x):
    """utils/yolov5_utils.py:220-227"""
    y = x.clone()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def non_max_suppression(prediction, conf_thres=0.4, iou_thres=0.35, max_det=300, max_wh=4096, max_nms=30000):
    """utils/yolov5_utils.py:124-218 with its defaults on the inference path (single label,
    class-aware, no merge).  prediction: (N, A, 5+nc) float32 tensor -> list of (n,6) tensors."""
    prediction = torch.as_tensor(prediction).clone()
    xc = prediction[..., 4] > conf_thres  # :136
    out = [torch.zeros((0, 6))] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]  # :152
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]  # :169
        box = xywh2xyxy(x[:, :4])  # :172
        conf, j = x[:, 5:].max(1, keepdim=True)  # :179
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]  # :180
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]  # :194-195
        c = x[:, 5:6] * max_wh  # :198
        boxes, scores = x[:, :4] + c, x[:, 4]
        i = torchvision.ops.nms(boxes, scores, iou_thres)  # :200
        if i.shape[0] > max_det:
            i = i[:max_det]
        out[xi] = x[i]
    return out


def connected_components_cv2(img_u8):
    """The call the reference effectively makes (utils/textmask.py:93,113,138):
    cv2.connectedComponentsWithStats(img) with the defaults connectivity=8, ltype=CV_32S."""
    import cv2
    n, labels, stats, centroids = cv2.connectedComponentsWithStats(img_u8)
    return n, labels, stats, centroids
