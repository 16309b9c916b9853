"""Annotation writer on top of the B200 detector: the on-disk formats of the reference's batch tool
`model2annotations` (inference.py:19-70) --

  <name>.txt        YOLO labels of the text blocks: "1 cx cy w h" per block, normalised, '\\n'-joined without a
                    trailing newline (imgproc_utils.py:22-29 `get_yololabel_strings`, :40-51 `xyxy2yolo`)
  line-<name>.txt   one text line per row, 8 integers x1 y1 .. x4 y4 (`np.savetxt(.., fmt='%d')`)
  <name>.json       `[blk.to_dict() ...]` through a numpy-aware JSON encoder (io_utils.py:16-28), optional
  <name>.png        the page re-encoded as PNG;  mask-<name>.png  the refined mask (io_utils.py:48-53 `imwrite`)

-- produced with refine_mode=REFINEMASK_ANNOTATION, keep_undetected_mask=True like the reference.  SURVEY 8f row f2.
"""
import glob
import json
import os
import os.path as osp
from pathlib import Path

import numpy as np

from .inference import REFINEMASK_ANNOTATION, TextDetector

IMG_EXT = (".bmp", ".jpg", ".png", ".jpeg")


class NumpyEncoder(json.JSONEncoder):
    """arrays -> lists, numpy scalars -> Python scalars (io_utils.py:16-28)."""

    def default(self, obj):
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, np.bool_):
            return bool(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.integer):
            return int(obj)
        return json.JSONEncoder.default(self, obj)


def find_all_imgs(img_dir, abs_path=False):
    """io_utils.py:30-41: every file of `img_dir` whose suffix is an image extension, in glob order."""
    out = []
    for fp in glob.glob(osp.join(img_dir, "*")):
        name = osp.basename(fp)
        if Path(name).suffix.lower() in IMG_EXT:
            out.append(fp if abs_path else name)
    return out


def imread(path, read_type=None):
    import cv2
    return cv2.imdecode(np.fromfile(path, dtype=np.uint8), cv2.IMREAD_COLOR if read_type is None else read_type)


def imwrite(img_path, img, ext=".png"):
    """io_utils.py:48-53: the suffix is REPLACED by `ext` (first occurrence of the suffix string in the path)."""
    import cv2
    suffix = Path(img_path).suffix
    img_path = img_path.replace(suffix, ext) if suffix != "" else img_path + ext
    cv2.imencode(ext, img)[1].tofile(img_path)


def xyxy2yolo(xyxy, w, h):
    """imgproc_utils.py:40-51: [x1,y1,x2,y2] -> normalised [cx,cy,w,h] (float64); None for no boxes."""
    if len(xyxy) == 0:
        return None
    a = np.array(xyxy)
    if a.ndim == 1:
        a = a[None]
    yolo = np.copy(a).astype(np.float64)
    yolo[:, [0, 2]] = yolo[:, [0, 2]] / w
    yolo[:, [1, 3]] = yolo[:, [1, 3]] / h
    yolo[:, [2, 3]] -= yolo[:, [0, 1]]
    yolo[:, [0, 1]] += yolo[:, [2, 3]] / 2
    return yolo


def get_yololabel_strings(clslist, labellist):
    """imgproc_utils.py:22-29"""
    rows = [str(int(c)) + " " + " ".join(str(e) for e in xywh) for c, xywh in zip(clslist, labellist)]
    return "\n".join(rows)


def write_annotations(save_dir, imgname, img, mask_refined, blk_list, save_json=False):
    """The per-page part of `model2annotations` (inference.py:33-70) for an already detected page."""
    im_h, im_w = img.shape[:2]
    imname = imgname.replace(Path(imgname).suffix, "")
    polys, blk_xyxy, blk_dicts = [], [], []
    for blk in blk_list:
        polys += blk.lines
        blk_xyxy.append(blk.xyxy)
        blk_dicts.append(blk.to_dict())
    yolo = xyxy2yolo(blk_xyxy, im_w, im_h)
    label = get_yololabel_strings([1] * len(yolo), yolo) if yolo is not None else ""
    with open(osp.join(save_dir, imname + ".txt"), "w", encoding="utf8") as f:
        f.write(label)
    if len(polys) != 0:
        np.savetxt(osp.join(save_dir, "line-" + imname + ".txt"), np.array(polys).reshape(-1, 8), fmt="%d")
    if save_json:
        with open(osp.join(save_dir, imname + ".json"), "w", encoding="utf8") as f:
            f.write(json.dumps(blk_dicts, ensure_ascii=False, cls=NumpyEncoder))
    imwrite(osp.join(save_dir, imgname), img)
    imwrite(osp.join(save_dir, "mask-" + imname + ".png"), mask_refined)


def model2annotations(model_path, img_dir_list, save_dir, save_json=False, detector=None):
    """`model2annotations(model_path, img_dir_list, save_dir, save_json)` of the reference (inference.py:19-70)."""
    if isinstance(img_dir_list, str):
        img_dir_list = [img_dir_list]
    det = detector if detector is not None else TextDetector(model_path=model_path, input_size=1024, act="leaky")
    os.makedirs(save_dir, exist_ok=True)
    try:
        imglist = []
        for d in img_dir_list:
            imglist += find_all_imgs(d, abs_path=True)
        for img_path in imglist:
            img = imread(img_path)
            _mask, mask_refined, blk_list = det(img, refine_mode=REFINEMASK_ANNOTATION, keep_undetected_mask=True)
            write_annotations(save_dir, osp.basename(img_path), img, mask_refined, blk_list, save_json)
    finally:
        if detector is None:
            det.close()
