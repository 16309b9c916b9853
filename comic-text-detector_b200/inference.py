"""Drop-in `TextDetector` over the B200 engine.

Same constructor and call signature as the reference's `inference.TextDetector`
(inference.py:116-178): `TextDetector(model_path, input_size=1024, device=..., half=False, nms_thresh=0.35,
conf_thresh=0.4, mask_thresh=0.3, act='leaky')` and
`detector(img, refine_mode=REFINEMASK_INPAINT, keep_undetected_mask=False) -> (mask, mask_refined, blk_list)`.

Everything with per-pixel or per-contour work runs in libctd_b200.so (network, NMS, mask u8, DB binarize,
connected components, contour boxes + scores, refine_mask); the host keeps what the reference does with
a handful of numbers per block (ratio scaling, `group_output`, window expansion) -- SURVEY section 7.
"""
from pathlib import Path
from typing import List

import numpy as np

from . import compiler
from .binding import Engine, PREC_FP16_TC, PREC_FP32_SIMT
from .textblock import TextBlock, group_output, overlap_area

REFINEMASK_INPAINT = 0
REFINEMASK_ANNOTATION = 1


def letterbox_geometry(shape_hw, new_shape=(1024, 1024)):
    """The size arithmetic of `letterbox(im, new_shape, auto=False)` (utils/imgproc_utils.py:86-117): aspect-preserving
    scale r, resized size (w, h) = round(size * r) (Python's round), bottom/right padding (dw, dh)."""
    r = min(new_shape[0] / shape_hw[0], new_shape[1] / shape_hw[1])
    new_unpad = int(round(shape_hw[1] * r)), int(round(shape_hw[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    return r, new_unpad, int(dw), int(dh)


def letterbox(im, new_shape=(1024, 1024)):
    """Host restatement of the reference's `letterbox` (kept for tests / tools; the detector itself resizes on the
    GPU through `Engine.forward_resized`, bit-exact with cv2.INTER_LINEAR)."""
    r, new_unpad, dw, dh = letterbox_geometry(im.shape[:2], new_shape)
    if im.shape[:2][::-1] != new_unpad:
        import cv2
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    if dw or dh:
        padded = np.zeros((new_shape[0], new_shape[1], 3), np.uint8)
        padded[:im.shape[0], :im.shape[1]] = im
        im = padded
    return im, (r, r), (dw, dh)


def expand_textwindow(img_size, xyxy, expand_r=8):
    """utils/imgproc_utils.py:151-161"""
    im_h, im_w = img_size[:2]
    x1, y1, x2, y2 = xyxy
    w, h = x2 - x1, y2 - y1
    pad = int(round((max(h, w) * 0.25 + min(h, w) * 0.75) / expand_r))
    return [max(0, x1 - pad), max(0, y1 - pad), min(im_w - 1, x2 + pad), min(im_h - 1, y2 + pad)]


class TextDetector:
    lang_list = ['eng', 'ja', 'unknown']
    langcls2idx = {'eng': 0, 'ja': 1, 'unknown': 2}

    def __init__(self, model_path, input_size=1024, device='cuda', half=False, nms_thresh=0.35, conf_thresh=0.4,
                 mask_thresh=0.3, act='leaky', precision=None, device_index=0):
        if isinstance(model_path, (str, Path)):
            import torch
            ckpt = torch.load(str(model_path), map_location='cpu')  # reference basemodel.py:212
        else:
            ckpt = model_path  # already a checkpoint dict
        if isinstance(input_size, int):
            input_size = (input_size, input_size)
        self.input_size = input_size
        self.device = device
        self.half = half
        self.conf_thresh = conf_thresh
        self.nms_thresh = nms_thresh
        self.backend = 'b200'
        self.program = compiler.compile_checkpoint(ckpt, head_act=act)
        if precision is None:
            precision = PREC_FP16_TC
        # DB threshold is hard-coded 0.3 in the reference (inference.py:139 ignores mask_thresh)
        self.net = Engine(self.program, device=device_index, precision=precision, max_batch=1, max_h=input_size[0],
                          max_w=input_size[1], conf_thresh=conf_thresh, nms_thresh=nms_thresh, db_thresh=0.3)

    def close(self):
        self.net.close()

    def __call__(self, img, refine_mode=REFINEMASK_INPAINT, keep_undetected_mask=False):
        eng = self.net
        # preprocess_img (inference.py:72-83): the BGR<->RGB double flip cancels, the net sees BGR.  The letterbox
        # resize + padding run on the GPU (cv2-exact INTER_LINEAR); net-sized pages skip the resize kernel.
        im_h, im_w = img.shape[:2]
        _r, new_unpad, dw, dh = letterbox_geometry((im_h, im_w), self.input_size)
        if (im_h, im_w) == tuple(self.input_size):
            eng.forward(np.ascontiguousarray(img)[None])
        else:
            eng.forward_resized(img, new_unpad[1], new_unpad[0], self.input_size[0], self.input_size[1])
        resize_ratio = (im_w / (self.input_size[0] - dw), im_h / (self.input_size[1] - dh))

        # postprocess_yolo (inference.py:101-114) on the GPU NMS rows
        det = eng.detections()[0]
        det[..., [0, 2]] = det[..., [0, 2]] * resize_ratio[0]
        det[..., [1, 3]] = det[..., [1, 3]] * resize_ratio[1]
        blks = (det[..., 0:4].astype(np.int32), det[..., 5].astype(np.int32), np.round(det[..., 4], 3))

        boxes, scores = eng.text_lines()                          # SegDetectorRepresenter (inference.py:158)
        keep = np.where(scores[0] > 0.6)                          # box_thresh (inference.py:159-161)
        lines = boxes[0][keep]

        # postprocess_mask + crop + cv2.resize back to the page (inference.py:85-99,164-168), all on the GPU
        mask = eng.mask_u8_resized(self.input_size[0] - dh, self.input_size[1] - dw, im_h, im_w)
        if lines.size == 0:
            lines = []
        else:
            lines = lines.astype(np.float64)
            lines[..., 0] *= resize_ratio[0]
            lines[..., 1] *= resize_ratio[1]
            lines = lines.astype(np.int32)
        blk_list = group_output(blks, lines, im_w, im_h, mask)
        mask_refined = self._refine(img, mask, blk_list, refine_mode)
        if keep_undetected_mask:
            mask_refined = self._refine_undetected(img, mask, mask_refined, blk_list, refine_mode)
        return mask, mask_refined, blk_list

    # ---- textmask.py:159-169 ---------------------------------------------------------------------
    def _refine(self, img, mask, blk_list: List[TextBlock], refine_mode):
        wins = np.array([expand_textwindow(img.shape, blk.xyxy, expand_r=16) for blk in blk_list], np.int32).reshape(-1, 4)
        h, w = img.shape[:2]
        if (h * w) % 4 == 0:
            return self.net.refine_mask(img, mask, wins, refine_mode)
        # the kernel wants h*w % 4 == 0: pad the columns with zeros (windows lie inside the page, so their
        # contents and therefore the result are unchanged) and crop the padding off again
        wp = (w + 3) // 4 * 4
        img_p = np.zeros((h, wp, 3), np.uint8)
        img_p[:, :w] = img
        mask_p = np.zeros((h, wp), np.uint8)
        mask_p[:, :w] = mask
        return np.ascontiguousarray(self.net.refine_mask(img_p, mask_p, wins, refine_mode)[:, :w])

    # ---- textmask.py:135-156 ---------------------------------------------------------------------
    def _refine_undetected(self, img, mask_pred, mask_refined, blk_list, refine_mode):
        mask_pred[np.where(mask_refined > 30)] = 0                 # in place, like the reference (App. D #13)
        pred_t = np.where(mask_pred > 30, 255, 0).astype(np.uint8)  # cv2.threshold(.., 30, 255, BINARY)
        n, labels, stats = self.net.connected_components(pred_t, stats_cap=int(pred_t.size // 4 + 2))
        valid = np.where(stats[:, -1] > 50)[0]
        seg_blks = []
        if len(valid) > 0:
            for li in valid[1:]:
                x, y, w, h, area = stats[li]
                bbox = [x, y, x + w, y + h]
                score = -1
                for blk in blk_list:
                    s = overlap_area(blk.xyxy, bbox)
                    if s > score:
                        score = s
                if score / w / h < 0.5:
                    seg_blks.append(TextBlock(bbox))
        if len(seg_blks) > 0:
            mask_refined = np.bitwise_or(mask_refined, self._refine(img, mask_pred, seg_blks, refine_mode))
        return mask_refined
