"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) in the build container.

    python -m oracle.make_golden

The fixtures are small (sub-sampled maps) and committed; they pin the oracle (tests/test_cpu_oracle.py)
and, on the GPU box where /root/reference does not exist, the CUDA path (tests/test_gpu_golden.py).
"""
import os
import tempfile

import numpy as np
import torch

from oracle import ref_shim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def net_golden(ns, smooth, name):
    ck = synth.make_checkpoint(0, smooth=smooth)
    f = tempfile.NamedTemporaryFile(suffix=".ckpt", delete=False).name
    torch.save(ck, f)
    net = ns.basemodel.TextDetBase(f, device="cpu", act="leaky")
    pages = np.stack([synth.structured_page(1000, 256, 256), synth.noise_page(1001, 256, 256)])
    xs = []
    for p in pages:  # the reference's own preprocessing (inference.py:72-83)
        x, ratio, dw, dh = ns.inference.preprocess_img(p, input_size=(256, 256), device="cpu")
        assert dw == 0 and dh == 0
        xs.append(x)
    x = torch.cat(xs, 0)
    with torch.no_grad():
        blks, mask, lines = net(x)
    det = ns.yolov5_utils.non_max_suppression(blks.clone(), 0.4, 0.35)
    np.savez_compressed(os.path.join(OUT, name), ckpt_seed=0, smooth=int(smooth), page_seeds=np.array([1000, 1001]),
                        blks_sub=blks[:, ::17].numpy(), mask_sub=mask[:, :, ::4, ::4].numpy(),
                        lines_sub=lines[:, :, ::4, ::4].numpy(),
                        mask_u8_sub=ns.inference.postprocess_mask(mask[0:1].clone())[::4, ::4],
                        det0=det[0].numpy(), det1=det[1].numpy())
    os.unlink(f)


def nms_golden(ns):
    rng = np.random.default_rng(5)
    rows = 3000
    p = np.zeros((1, rows, 7), np.float32)
    p[0, :, 0:2] = rng.uniform(0, 1024, (rows, 2))
    p[0, :, 2:4] = rng.uniform(4, 200, (rows, 2))
    p[0, :, 4] = rng.uniform(0, 1, rows) ** 2
    p[0, :, 5:] = rng.uniform(0, 1, (rows, 2))
    p[0, 100:200] = p[0, 0:100]
    p[0, 100:200, :4] += rng.normal(0, 2, (100, 4)).astype(np.float32)
    out = ns.yolov5_utils.non_max_suppression(torch.from_numpy(p.copy()), 0.4, 0.35)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "nms_ref.npz"), pred=p[0], det=out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load()
    net_golden(ns, True, "net_smooth_256.npz")
    net_golden(ns, False, "net_rough_256.npz")
    nms_golden(ns)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
