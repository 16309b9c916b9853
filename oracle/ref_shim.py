"""TEST INFRASTRUCTURE ONLY (oracle): import the UNMODIFIED reference from /root/reference.

The reference (dmMaze/comic-text-detector @ 440b978) is pure Python.  It imports
here with three shims, none of which touch its arithmetic:

  * ``torchsummary`` is absent -> stub module with a no-op ``summary`` (basemodel.py:13)
  * ``np.bool8`` / ``np.float_`` were removed in numpy 2 (utils/io_utils.py:11-12)
  * ``pyclipper`` / ``shapely`` are absent and un-vendored (utils/db_utils.py:3-4,
    utils/textblock.py:3) -> replaced by the restatements in ``oracle/geom_ref.py``
    (parity UNPINNED for those two third-party calls).

/root/reference exists only in the build container, never on the GPU box, so this
module is used exclusively by ``oracle/make_golden.py`` (fixture generation) and by
``-m "not gpu"`` tests that validate the restatements in ``oracle/`` against it.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("CTD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "basemodel.py"))


_loaded = {}


def load():
    """Returns a namespace with the reference modules (basemodel, yolo, yolov5_utils, ...)."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import numpy as np
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if "torchsummary" not in sys.modules:
        m = types.ModuleType("torchsummary")
        m.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = m
    # third-party geometry the reference imports but the image lacks
    from oracle import geom_ref
    if "pyclipper" not in sys.modules:
        sys.modules["pyclipper"] = geom_ref.make_pyclipper_module()
    if "shapely" not in sys.modules:
        shp, geo = geom_ref.make_shapely_modules()
        sys.modules["shapely"] = shp
        sys.modules["shapely.geometry"] = geo
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.environ.setdefault("WANDB_MODE", "disabled")
    import importlib
    ns = types.SimpleNamespace()
    ns.basemodel = importlib.import_module("basemodel")
    ns.yolo = importlib.import_module("models.yolov5.yolo")
    ns.common = importlib.import_module("models.yolov5.common")
    ns.yolov5_utils = importlib.import_module("utils.yolov5_utils")
    ns.db_utils = importlib.import_module("utils.db_utils")
    ns.textblock = importlib.import_module("utils.textblock")
    ns.textmask = importlib.import_module("utils.textmask")
    ns.imgproc_utils = importlib.import_module("utils.imgproc_utils")
    ns.inference = importlib.import_module("inference")
    _loaded["ns"] = ns
    return ns
