#!/usr/bin/env python
"""bench.py -- pages/sec of the comic-text-detector hot path on N B200s (driver contract).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--batch B]

A "step" = one pass of the hot path (backbone + seg head + DB head + Detect decode + NMS +
mask u8 + DB threshold + connected components + text-line boxes/scores) over one batch of B synthetic 1024x1024 pages
per GPU (BASELINE.json configs[2]/[3]: batch 16 per GPU, fp16 tcgen05 path).

* value      : whole-job pages/s with the pages already resident in HBM (device timed, CUDA events
               on the engine stream, max over ranks).  --engines E (default 2) workspaces per GPU: consecutive
               steps alternate between them, so E CUDA graphs are in flight; a step is always one full batch.
* e2e        : same metric through the C-ABI with HOST (pinned) page buffers (ctd_submit / ctd_collect): H2D of
               the pages and D2H of the results (mask u8 + detections + text-line boxes/scores + counts) of
               EVERY step inside the timed region, copies overlapped with the neighbouring steps' compute;
               e2e.sync_value = the blocking ctd_forward + ctd_get_* sequence on one engine.
* roofline   : tensor roofline of the tcgen05 convolution kernels (conv_tc / conv_halo / conv_hs): algorithmic
               conv FLOPs of the tensor-core layers / their summed device time (per-op CUDA events, measured
               live here, serial order), against MEASURED_PEAKS.json's sustained bf16 GEMM rate; traffic = DRAM
               bytes of the same launches from the committed ncu capture.
* cpu_baseline / --impl reference: the oracle restatement of the reference's CPU path
               (oracle/net_ref.py + oracle/postproc_ref.py: torch CPU fp32 + torchvision + cv2, i.e.
               the reference's own library calls) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_PAGE_1024 = 191.414  # BASELINE.md section 2 (2*MAC over the 115 conv/deconv layers)


def conv_flops(prog, n, h, w):
    """Algorithmic FLOPs of each op of the program at batch n (0 for non-conv ops)."""
    out = []
    for o in prog.ops:
        kind = o["kind"]
        if kind == 0:  # stem: 6x6 s2 conv 3 -> cout (algorithmic FLOPs, not the zero-padded tensor-core K)
            out.append(2.0 * (h // 2) * (w // 2) * n * 108 * o["cout"])
            continue
        if kind not in (1, 2, 6, 7):
            out.append(0.0)
            continue
        down = prog.bufs[o["src_buf"][0]][1]
        px = (h // down) * (w // down) * n
        cin = sum(o["src_c"][: o["n_src"]])
        if kind == 7:   # seg tail: ConvT 4x4 s2, C -> 1 (16 taps per input pixel)
            out.append(2.0 * px * 16 * cin)
        elif kind == 2:
            out.append(2.0 * px * 16 * cin * o["cout"])
        else:
            k, s = o["ksize"], o["stride"]
            out.append(2.0 * (px // (s * s)) * k * k * cin * o["cout"])
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_pipeline_factory(h, w):
    """The reference's CPU path for the same stages, via the oracle restatement."""
    import torch
    from oracle import synth, postproc_ref
    from oracle.net_ref import RefNet
    ck = synth.make_checkpoint(0, smooth=True)
    net = RefNet(ck)

    def run(page_u8):
        x = torch.from_numpy(np.ascontiguousarray(page_u8.transpose(2, 0, 1))[None].astype(np.float32) / 255)
        blks, mask, lines = net(x)
        det = postproc_ref.non_max_suppression(blks, 0.4, 0.35)[0]
        m8 = (mask[0, 0].numpy() * 255).astype(np.uint8)
        bitmap = (lines[0, 0].numpy() > 0.3).astype(np.uint8)
        n, labels, stats, _ = postproc_ref.connected_components_cv2(bitmap)
        boxes, scores = postproc_ref.seg_represent(lines[0, 0].numpy(), 0.3)
        return det, m8, n, boxes, scores
    return run


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    cores = min(os.cpu_count(), args.cpu_threads)
    torch.set_num_threads(cores)
    run = cpu_pipeline_factory(1024, 1024)
    pages = [synth.structured_page(1000 + i) for i in range(max(1, args.cpu_pages))]
    for _ in range(args.warmup):
        run(pages[0])
    t0 = time.perf_counter()
    for k in range(args.steps):
        for p in pages:
            run(p)
    dt = time.perf_counter() - t0
    npages = args.steps * len(pages)
    val = npages / dt
    print(json.dumps({
        "impl": "reference", "metric": "pages/sec @1024x1024 synthetic", "value": val, "unit": "pages/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference CPU path (oracle port: torch CPU fp32 forward + torchvision NMS + cv2 CC + SegDetectorRepresenter), "
                               "%d page(s) of 1024x1024 per step, all host threads" % len(pages)},
        "cpu_baseline": {"value": val, "unit": "pages/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d structured synthetic 1024x1024 page(s)" % (args.steps, len(pages))},
        "e2e": {"value": val, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=16, help="pages per GPU per step")
    ap.add_argument("--cpu-pages", type=int, default=1, help="pages per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engines", type=int, default=2, help="workspaces per GPU; consecutive batches alternate between them")
    ap.add_argument("--api-pages", type=int, default=8, help="pages timed through the TextDetector Python API (0 = skip)")
    ap.add_argument("--cpu-threads", type=int, default=16,
                    help="torch intra-op threads of the CPU arm (measured on the B200 host: 16 threads 0.25 s/forward, "
                         "64 threads 0.48 s, 128 threads 33 s -- more threads only hurt)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import ctd_b200
    from oracle import synth  # synthetic checkpoint + pages only (no oracle compute on this path)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, H, W = args.batch, 1024, 1024

    ck = synth.make_checkpoint(0, smooth=True)
    prog = ctd_b200.compiler.compile_checkpoint(ck)
    eng = ctd_b200.Engine(prog, device=local, max_batch=B, max_h=H, max_w=W, use_graph=True)
    pages = np.stack([synth.structured_page(1000 + rank * B + i) for i in range(B)])
    host_pages = torch.from_numpy(pages).pin_memory()
    dev_pages = host_pages.cuda()
    torch.cuda.synchronize()
    # pinned result buffers for the e2e leg
    out_mask = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    out_det = torch.empty((B, 300, 6), dtype=torch.float32).pin_memory()
    out_cnt = torch.empty((B,), dtype=torch.int32).pin_memory()
    out_nl = torch.empty((B,), dtype=torch.int32).pin_memory()
    out_lb = torch.empty((B, 1000, 4, 2), dtype=torch.int16).pin_memory()
    out_ls = torch.empty((B, 1000), dtype=torch.float32).pin_memory()
    out_lc = torch.empty((B,), dtype=torch.int32).pin_memory()
    lib, hnd = eng.lib, eng.h
    import ctypes as C

    # Two workspaces (engines) per GPU: consecutive batches alternate between them, so two CUDA graphs are in
    # flight and the kernels of batch i+1 fill the tails / dependency gaps of batch i.  Every step is still one full
    # forward of B pages; the timer (engine 0's stream) is closed after ctd_join has pulled in the other streams.
    n_eng = max(1, args.engines)
    engs = [eng] + [ctd_b200.Engine(prog, device=local, max_batch=B, max_h=H, max_w=W, use_graph=True)
                    for _ in range(n_eng - 1)]
    ctr = {"res": 0, "e2e": 0}

    def step_resident():
        e = engs[ctr["res"] % n_eng]
        ctr["res"] += 1
        e.forward_device(dev_pages.data_ptr(), B, H, W)
        return e

    def step_e2e_sync():
        # one blocking call after the other (ctd_forward + ctd_get_*), one engine, nothing overlapped
        eng._ck(lib.ctd_forward(hnd, C.c_void_p(host_pages.data_ptr()), B, H, W, 0))
        eng.shape = (B, H, W)
        eng._ck(lib.ctd_get_mask_u8(hnd, C.c_void_p(out_mask.data_ptr())))
        eng._ck(lib.ctd_get_detections(hnd, C.c_void_p(out_det.data_ptr()), C.c_void_p(out_cnt.data_ptr())))
        eng._ck(lib.ctd_get_db_components(hnd, None, None, C.c_void_p(out_nl.data_ptr())))
        eng._ck(lib.ctd_get_text_lines(hnd, C.c_void_p(out_lb.data_ptr()), C.c_void_p(out_ls.data_ptr()),
                                       C.c_void_p(out_lc.data_ptr())))

    # pipelined host path (ctd_submit / ctd_collect): every step still copies its own pages H2D from pinned
    # memory and its own result arena D2H, but step i+1's upload and step i-1's download run under step i
    res_bytes = eng.results_bytes()
    out_arena = [[torch.empty((res_bytes,), dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(n_eng)]
    pending = []

    def step_e2e():
        k = ctr["e2e"]
        ctr["e2e"] += 1
        ei, slot = k % n_eng, (k // n_eng) & 1
        if len(pending) == 2 * n_eng:
            pe, ps = pending.pop(0)
            engs[pe].collect(ps)
        engs[ei].submit(slot, host_pages.data_ptr(), B, H, W, out_arena[ei][slot].data_ptr())
        pending.append((ei, slot))

    def drain_e2e():
        while pending:
            pe, ps = pending.pop(0)
            engs[pe].collect(ps)

    def join_all():
        for e in engs[1:]:
            eng.join(e)

    step_main = step_resident
    if world > 1:
        # single NCCL gather of each rank's result arena (mask u8 | detections | counts) to rank 0 over
        # NVLink, issued on the stream of the engine that produced it so the device timer covers it
        class _DevArr:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
        gat = {}
        for e in engs:
            e.forward_device(dev_pages.data_ptr(), B, H, W)
            o = e.device_outputs()
            res_t = torch.as_tensor(_DevArr(o.results_base, o.results_bytes), device="cuda")
            glist = [torch.empty_like(res_t) for _ in range(world)] if rank == 0 else None
            gat[id(e)] = (res_t, glist, torch.cuda.ExternalStream(o.stream))

        def step_main():
            e = step_resident()
            res_t, glist, ext = gat[id(e)]
            with torch.cuda.stream(ext):
                dist.gather(res_t, gather_list=glist, dst=0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        eng.timer_start()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()  # host-blocks until the last D2H landed, so the stop event is recorded after it
        join_all()   # the other engines' streams become dependencies of the timer stream
        ms = eng.timer_stop()
        barrier()
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(3, args.warmup) * n_eng):
        step_main()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step_main, args.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(3 * n_eng):
        step_e2e()
    drain_e2e()
    ms_e2e = timed(step_e2e, args.steps, drain_e2e)
    step_e2e_sync()
    ms_e2e_sync = timed(step_e2e_sync, args.steps)

    # per-op device times of one forward -> roofline of the tensor-core conv kernel
    op_ms, nms_ms, ccl_ms = eng.profile_forward(dev_ptr=dev_pages.data_ptr(), shape=(B, H, W))
    op_ms2, _, _ = eng.profile_forward(dev_ptr=dev_pages.data_ptr(), shape=(B, H, W))
    op_ms = np.minimum(op_ms, op_ms2)
    fl = conv_flops(prog, B, H, W)
    tc_idx = [i for i, o in enumerate(prog.ops) if o["kind"] in (0, 1, 2, 6, 7)]
    tc_ms = float(sum(op_ms[i] for i in tc_idx))
    tc_flops = float(sum(fl[i] for i in tc_idx))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    # DRAM bytes of the same launches from the committed ncu capture (profiles/): not measurable live
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_traffic.json")))
        if int(tj.get("batch", 0)) == B:
            traffic = float(tj["dram_bytes"])
            traffic_src = "profiles/r01_conv_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d conv launches of one step)" % int(tj["launches"])
    except Exception:
        pass

    if rank == 0:
        total_pages = B * world * args.steps
        value = total_pages / (ms * 1e-3)
        e2e_val = total_pages / (ms_e2e * 1e-3)
        line = {
            "metric": "pages/sec @1024x1024 synthetic", "value": value, "unit": "pages/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: 1024x1024 pages, batch %d per GPU, fp16 tcgen05 path, full device "
                                   "pipeline (backbone + seg head + DB head + Detect/NMS + mask u8 + DB threshold + CCL + contour boxes/scores)" % B,
                       "pages_per_gpu_per_step": B, "page": [H, W], "checkpoint": "synthetic seed 0 (oracle/synth.py)",
                       "l2": "activations per step (~%.1f GB) exceed the 126 MB L2; no explicit flush" % (
                           sum(c * (H // d) * (W // d) for c, d in prog.bufs) * 2 * B / 1e9),
                       "cuda_graph": True,
                       "engines_per_gpu": n_eng,
                       "in_flight": "%d batches per GPU (one CUDA graph each, alternating workspaces)" % n_eng,
                       "multi_gpu": "pages sharded B per rank; one NCCL gather of each rank's result arena to rank 0 per step" if world > 1 else "single GPU"},
            "gpu_launches": eng.last_launch_count() * args.steps,
            "clocks": clocks,
            "conv_roofline_frac_of_nominal": value / world * GFLOP_PER_PAGE_1024 * 1e9 / 2.25e15,
            "e2e": {"value": e2e_val, "unit": "pages/s", "h2d_bytes_per_step": int(B * H * W * 3),
                    "d2h_bytes_per_step": int(res_bytes),
                    "mode": "ctd_submit/ctd_collect on %d engine(s) per GPU, two batches in flight per engine (copies under compute), pinned host buffers" % n_eng,
                    "sync_value": total_pages / (ms_e2e_sync * 1e-3),
                    "sync_mode": "ctd_forward + ctd_get_* blocking, nothing overlapped"},
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel + conv_halo_kernel, the tcgen05 implicit-GEMM convolutions (%d launches per step)" % len(tc_idx), "achieved": achieved,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_unit": "bytes/step", "traffic_source": traffic_src,
                         "peak_source": peak_src, "flops_per_step": tc_flops, "ms_per_step": tc_ms,
                         "share_of_step": tc_ms / float(op_ms.sum() + nms_ms + ccl_ms)},
            "stage_ms": {"conv_tc": tc_ms, "other_ops": float(op_ms.sum()) - tc_ms, "nms": nms_ms, "ccl_and_line_boxes": ccl_ms},
        }
        if world == 1 and args.api_pages > 0:
            # the drop-in Python API, one page per call (TextDetector.__call__: H2D, all GPU stages, host
            # group_output, GPU refine_mask, D2H of masks): what a caller of the reference's interface sees
            det = ctd_b200.TextDetector(ck, input_size=1024, act="leaky")
            det(pages[0].copy())
            t0 = time.perf_counter()
            nblk = 0
            for i in range(args.api_pages):
                _m, _mr, _bl = det(pages[i % B].copy())
                nblk += len(_bl)
            dt = time.perf_counter() - t0
            det.close()
            line["api_e2e"] = {"value": args.api_pages / dt, "unit": "pages/s", "pages": args.api_pages,
                               "blocks_per_page": nblk / args.api_pages,
                               "what": "TextDetector.__call__ per page, single stream, incl. host group_output"}
        if not args.no_cpu_baseline and world == 1:
            cores = min(os.cpu_count(), args.cpu_threads)
            torch.set_num_threads(cores)
            run = cpu_pipeline_factory(H, W)
            run(pages[0])
            t0 = time.perf_counter()
            ncpu = 1
            for i in range(ncpu):
                run(pages[i % B])
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ncpu / dt, "unit": "pages/s", "cores": cores, "kind": "port",
                                    "sample": "%d structured synthetic 1024x1024 pages, same stages, oracle port of the "
                                              "reference CPU path (torch fp32 + torchvision NMS + cv2 CC + SegDetectorRepresenter)" % ncpu}
        print(json.dumps(line))
    for e in engs:
        e.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
