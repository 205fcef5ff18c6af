#!/usr/bin/env python
"""Benchmark of the SIFT hot path (detect + describe) on B200 -- BASELINE.json's metric
"4K SIFT images/sec (detect+describe)" on config "synthetic 3840x2160 (4K) batch of 256 images
sharded by image across 1/2/4/8 B200".

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference ...                            the reference's CPU path (oracle/_ref)

One step = one pass of detect+describe over the whole batch (256 images, image i on rank i mod N:
total work fixed => "strong" scaling).  `value` is timed with the images already in HBM and the
results left in HBM; `e2e` goes through the host-buffer C-ABI calls (pinned host pixels in,
168-byte keypoint records out) with both copies inside the timed region.  Prints ONE JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
METRIC = "4K SIFT images/sec (detect+describe)"
HBM_BYTES_PER_INPUT_PIXEL = 321.0  # compulsory stage-boundary traffic, doubling on


# ------------------------------------------------------------------------------------------
# synthetic data: generator "D"  -- sum of unit-variance Gaussian-filtered noise
# fields at sigma 2,4,8,16,32 with wrap-around, mapped to [0,255] u8.  Built on the GPU with FFTs
# (circular convolution == scipy's mode="wrap"); seed 1234 + image index.
# ------------------------------------------------------------------------------------------
def synth_image_gpu(h, w, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def response(n, s, half):
        # scipy.ndimage.gaussian_filter's kernel: radius int(4 sigma + 0.5), normalised, wrapped
        r = int(4.0 * s + 0.5)
        x = torch.arange(-r, r + 1, device=device, dtype=torch.float64)
        k = torch.exp(-0.5 * x * x / (s * s))
        k = k / k.sum()
        line = torch.zeros(n, device=device, dtype=torch.float64)
        line.index_add_(0, (x.long() % n), k)
        return (torch.fft.rfft(line) if half else torch.fft.fft(line)).real.float()

    acc = torch.zeros((h, w), device=device)
    for s in (2.0, 4.0, 8.0, 16.0, 32.0):
        f = torch.fft.rfft2(torch.randn((h, w), generator=g, device=device, dtype=torch.float32))  # fresh field
        n = torch.fft.irfft2(f * response(h, s, False).view(-1, 1) * response(w, s, True).view(1, -1), s=(h, w))
        acc += n / n.std()
    acc = (acc - acc.min()) / (acc.max() - acc.min()) * 255.0
    return torch.round(acc).to(torch.uint8).contiguous()


def pyramid_algorithmic_bytes(w, h, doubled=True):
    """Compulsory HBM bytes of the Gaussian/DoG pyramid stage for one image:
    per octave read the base (4 B/px), write G1..G3 (12), write D0..D4 (20), write the next base."""
    bw, bh = (2 * w, 2 * h) if doubled else (w, h)
    octaves = int(math.floor(math.log2(min(bw, bh) // 3)))
    total = 0
    for o in range(octaves):
        p = bw * bh
        total += 36 * p
        bw, bh = bw // 2, bh // 2
        if o + 1 < octaves:
            total += 4 * bw * bh
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return None
        time.sleep(0.25)
        self.p.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 - 0.1 <= t <= t1 + 0.3:
                    sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation (oracle/_ref, the real
# reference compiled from its sources; "port" = oracle/liboracle.so if that is absent)
# ------------------------------------------------------------------------------------------
def _cpu_worker_init():
    global _W
    from oracle import oracle as O
    d = tempfile.mkdtemp(prefix="siftref_")
    os.chdir(d)  # the reference's public entry point writes ./keypoints.png (sift.cpp:765-768)
    _W = {"O": O, "kind": "reference" if O.have_ref() else "port"}
    if _W["kind"] == "reference":
        O.ref()
    else:
        O.port()


def _cpu_detect(img, want_keypoints=False):
    O = _W["O"]
    t = time.perf_counter()
    if _W["kind"] == "reference":
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # silence the reference's per-stage std::cout chatter
        try:
            k = O.ref_detect_public(img)
        finally:
            os.dup2(saved, 1)
            os.close(devnull); os.close(saved)
    else:
        k = O.Run(O.port(), img, keep_pyramid=False).keypoints(2)
    dt = time.perf_counter() - t
    return (len(k), dt, k) if want_keypoints else (len(k), dt)


def cpu_kind():
    from oracle import oracle as O
    return "reference" if O.have_ref() else "port"


def host_synth_crop(h, w, seed):
    from oracle import oracle as O
    return O.synth_image(h, w, seed=seed)


# sample sizes of the reference arm: (height, width, seconds one core needs, measured with oracle/_ref on this pool's
# hosts).  The reference is single-threaded (no <thread>, OpenMP or SIMD in src/), so "all the host threads it can
# use" = one independent process per core, each on its own image.
REF_SAMPLES = [(2160, 3840, 130.0), (1080, 1920, 24.0), (540, 960, 5.5)]


def run_reference_arm(args):
    """bench.py --impl reference: every host core runs the reference's detect_keypoints_and_descriptors
    (sift.cpp:712-776, compiled from the reference's sources) on its own generator-D image per step.  The sample is
    the largest of 4K / 1080p / 960x540 that keeps the whole --steps/--warmup run within ~4 minutes; smaller
    samples are scaled to 4K images/s by pixel count, which FLATTERS the CPU (4K costs 4.9x a 1080p image for 4x
    the pixels) -- same_config is true only for the 4K sample.  The line also carries one 1080p
    timing on one core, so the scaling cpu_baseline uses is visible in the same run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    except (ValueError, OSError):
        mem_gb = 64.0
    n_steps = args.warmup + args.steps
    budget_s = 240.0
    sh, sw, est = REF_SAMPLES[-1]
    for h, w, t in REF_SAMPLES:
        need_gb = cores * 4.5 * (h * w) / (H4K * W4K)      # the FP64 pyramid of a 4K image is ~4.3 GB
        if t * n_steps <= budget_s and need_gb < 0.7 * mem_gb:
            sh, sw, est = h, w, t
            break
    frac = (sh * sw) / (H4K * W4K)
    imgs = [host_synth_crop(sh, sw, 1234 + i) for i in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        times = []
        for step in range(n_steps):
            t = time.perf_counter()
            res = pool.map(_cpu_detect, imgs, chunksize=1)
            dt = time.perf_counter() - t
            if step >= args.warmup:
                times.append(dt)
        one_1080p = None
        if (sh, sw) != (1080, 1920):
            n1, t1 = pool.apply(_cpu_detect, (host_synth_crop(1080, 1920, 1234),))
            one_1080p = {"seconds_one_core": t1, "keypoints": n1,
                         "images_per_s_4k_equiv_all_cores": cores * 0.25 / t1,
                         "note": "one 1920x1080 image on one otherwise idle core, scaled by pixel count and by the core "
                                 "count: the scaling cpu_baseline (1 core) uses"}
    total = sum(times)
    value = args.steps * cores * frac / total
    kind = cpu_kind()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 3840x2160 (4K) batch of 256 images, detect+describe",
                   "sample": f"{cores} x {sw}x{sh} generator-D images per step (one per core, {frac:.4f} of a 4K image each)"
                             + ("" if frac == 1.0 else ", scaled to 4K images/s by pixel count (flatters the CPU: "
                                                       "cost grows faster than pixels)"),
                   "same_config": frac == 1.0,
                   "keypoints_per_sample": float(np.mean([r[0] for r in res])),
                   "seconds_per_sample_per_core": total / args.steps,
                   "one_1080p": one_1080p, "host_memory_gb": mem_gb},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} processes x one {sw}x{sh} image per step; reference "
                                   f"{'copy-free build of /root/reference/src (bit-identical output)' if kind == 'reference' else 'CPU port oracle/sift_oracle.cpp'}"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def config1_pair(device):
    """BASELINE.json config 1 (stitching/image1.jpg + image2.jpg, stb-decoded pixels in tests/golden): this repo's
    detect x2 + match through the host-buffer C ABI, beside the reference on the same pair on one core
    (copy-free build, bit-identical output) and the recorded wall time of the reference AS SHIPPED
    (profiles/r2_config1_asshipped.json: the unmodified ./sift needs minutes because it deep-copies whole octaves per
    extremum, sift.cpp:346 -- too long to repeat inside every benchmark run)."""
    import multiprocessing as mp
    from PIL import Image
    import sift_project_b200 as S
    px = [np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", n + ".png"))) for n in ("image1", "image2")]
    with S.SiftContext(px[0].shape[1], px[0].shape[0], device=device) as c:
        for _ in range(2):
            t = time.perf_counter()
            k = [c.detect(p) for p in px]
            ia, ib, d = c.match(k[0]["desc"], k[1]["desc"])
            gpu_s = time.perf_counter() - t
    ctx = mp.get_context("spawn")
    with ctx.Pool(1, initializer=_cpu_worker_init) as pool:
        res = [pool.apply(_cpu_detect, (p.astype(np.float64), True)) for p in px]
        t = time.perf_counter()
        m = pool.apply(_cpu_match, (res[0][2]["desc"], res[1][2]["desc"]))
        cpu_match_s = time.perf_counter() - t
    shipped = None
    try:
        shipped = json.load(open(os.path.join(ROOT, "profiles", "r2_config1_asshipped.json")))
    except Exception:
        pass
    return {"workload": "stitching/image1.jpg + image2.jpg (755x499 RGB): detect x2 + match, host buffers in and out",
            "gpu_ms": 1e3 * gpu_s, "keypoints": [len(x) for x in k], "matches": len(ia),
            "cpu_copy_free_s": res[0][1] + res[1][1] + cpu_match_s, "cpu_keypoints": [res[0][0], res[1][0]],
            "cpu_matches": m, "cpu_cores": 1, "reference_as_shipped": shipped}


def _cpu_match(a, b):
    O = _W["O"]
    lib = O.ref() if _W["kind"] == "reference" else O.port()
    return len(O.match(lib, a, b)[0])


def _cpu_match_timed(n):
    O = _W["O"]
    lib = O.ref() if _W["kind"] == "reference" else O.port()
    a, b = O.synth_descriptors(n, seed=1234), O.synth_descriptors(n, seed=1235)
    t = time.perf_counter()
    O.match(lib, a, b)
    return time.perf_counter() - t


def cpu_match_baseline(n=4000):
    """The reference's match_keypoints loop (sift.cpp:789-812) on one core: GFLOP/s by the same 2*N1*N2*128 count."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(1, initializer=_cpu_worker_init) as pool:
        dt = pool.apply(_cpu_match_timed, (n,))
    return {"value": 2.0 * n * n * 128 / dt / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": cpu_kind(),
            "sample": f"{n} x {n} config-5 style descriptors, {dt:.2f} s"}


def cpu_baseline_sample(img_u8_host, device):
    """rank 0, N=1: the reference detect on one 1920x1080 crop (1/4 of a 4K image), 1 thread -- and, since the
    reference's keypoints of that crop are now at hand, the parity of the GPU path on the BENCHMARK's own pixels
    (the torch-FFT generator, not the scipy one of the tests)."""
    import multiprocessing as mp
    import sift_project_b200 as S
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity as P
    crop = np.ascontiguousarray(img_u8_host[:1080, :1920])
    ctx = mp.get_context("spawn")
    with ctx.Pool(1, initializer=_cpu_worker_init) as pool:
        n, dt, want = pool.apply(_cpu_detect, (crop, True))
    kind = cpu_kind()
    cpu = {"value": 0.25 / dt, "unit": "images/s", "cores": 1, "kind": kind,
           "sample": f"one 1920x1080 crop (1/4 of the pixels) of batch image 0, {n} keypoints, {dt:.1f} s, "
                     f"scaled by pixel count; "
                     f"{'reference sources compiled copy-free (oracle/_ref)' if kind == 'reference' else 'oracle port'}"}
    with S.SiftContext(1920, 1080, device=device) as c:
        got = c.detect(crop)
        stage = c.describe_given(want)      # the descriptor stage alone, on the reference's own keypoints
    rec, prec, gi, wi = P.recall_precision(got, want)
    rep = P.descriptor_report(got, want, gi, wi)
    d = np.abs(got["desc"][gi].astype(np.int16) - want["desc"][wi].astype(np.int16)).max(1) if len(gi) else np.zeros(0)
    g, w_ = got[gi][d > 1], want[wi][d > 1]
    same = (g["x"] == w_["x"]) & (g["y"] == w_["y"]) & (g["size"] == w_["size"]) & (g["pori"] == w_["pori"])
    parity = {"image": "1920x1080 crop of batch image 0 (this benchmark's generator), vs the reference run timed for cpu_baseline",
              "keypoints_gpu": len(got), "keypoints_ref": len(want), "recall": rec, "precision": prec,
              "pos_tol_px": P.POS_TOL, "size_rtol": P.SIZE_RTOL,
              "desc_frac_within_1": rep["frac_le1"], "desc_frac_exact": rep["frac_exact"], "desc_max": rep["max"],
              "n_outliers": int((d > 1).sum()), "outliers_with_identical_keypoint": int(same.sum()),
              "desc_stage_max": int(np.abs(stage["desc"].astype(np.int16) - want["desc"].astype(np.int16)).max())}
    return cpu, parity


# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import sift_project_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (no "NCCL version" banner)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    batch = args.images
    W, H = args.width, args.height
    mine = list(range(rank, batch, world))
    d_imgs = [synth_image_gpu(H, W, 1234 + i, dev) for i in mine]
    torch.cuda.synchronize()
    n_ctx = args.contexts
    ctxs = [S.SiftContext(W, H, device=local) for _ in range(n_ctx)]
    streams = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]
    main = torch.cuda.current_stream()

    # one verification pass: keypoint counts per image, no overflow
    counts = []
    for k, img in enumerate(d_imgs[: min(len(d_imgs), 4)]):
        ctxs[0].detect_enqueue(img, W, H)
        counts.append(ctxs[0].detect_finish())
    kp_mean = float(np.mean(counts)) if counts else 0.0

    def one_step_device():
        for k, img in enumerate(d_imgs):
            ctxs[k % n_ctx].detect_enqueue(img, W, H)

    def join_streams():
        for s in streams:
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)

    # ---- value: device-resident inputs and outputs ----
    for _ in range(args.warmup):
        one_step_device()
    join_streams()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = sum(c.launches for c in ctxs)
    t_wall0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(main)
    for s in streams:
        s.wait_event(start)
    for _ in range(args.steps):
        one_step_device()
    join_streams()
    end.record(main)
    barrier()
    t_wall1 = time.time()
    ms_dev = reduce_max(start.elapsed_time(end))
    launches = reduce_sum(sum(c.launches for c in ctxs) - l0)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    for c in ctxs:
        c.detect_finish()
    value = batch * args.steps / (ms_dev * 1e-3)

    # ---- e2e: pinned host pixels in, host keypoint records out, through the C-ABI ----
    h_imgs = [img.cpu().pin_memory() for img in d_imgs]
    cap = int(max(counts) * 1.5) + 4096 if counts else 65536
    outs = [np.zeros(cap, dtype=S.KP_DTYPE) for _ in range(n_ctx)]
    for o in outs:
        torch.cuda.cudart().cudaHostRegister(o.ctypes.data, o.nbytes, 0)

    def one_step_e2e():
        d2h = 0
        pending = [False] * n_ctx
        for k, img in enumerate(h_imgs):
            j = k % n_ctx
            if pending[j]:
                d2h += ctxs[j].result_copy(outs[j]) * 168
            ctxs[j].detect_enqueue(img, W, H)
            pending[j] = True
        for j in range(n_ctx):
            if pending[j]:
                d2h += ctxs[j].result_copy(outs[j]) * 168
        return d2h

    for _ in range(max(1, args.warmup // 2)):
        one_step_e2e()
    barrier()
    t0 = time.perf_counter()
    d2h_bytes = 0
    for _ in range(args.steps):
        d2h_bytes += one_step_e2e()
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_value = batch * args.steps / e2e_s
    h2d_step = reduce_sum(len(h_imgs) * W * H)
    d2h_step = reduce_sum(d2h_bytes / args.steps)

    # ---- the same end-to-end path for RGB input (what image_io.cpp:20-35 yields for a photo): 3 bytes per pixel
    # cross PCIe and the fused input kernel converts to gray on the GPU (rank 0, one GPU, a smaller batch) ----
    e2e_rgb = None
    if rank == 0 and world == 1:
        try:
            n_rgb = min(len(d_imgs), 32)
            h_rgb = []
            for k in range(n_rgb):
                g = d_imgs[k]
                rgb = torch.stack([g, torch.roll(g, 3, 1), torch.roll(g, 5, 0)], dim=-1).contiguous()
                h_rgb.append(rgb.cpu().pin_memory())
            del rgb

            def one_step_rgb():
                pending = [False] * n_ctx
                for k, img in enumerate(h_rgb):
                    j = k % n_ctx
                    if pending[j]:
                        ctxs[j].result_copy(outs[j])
                    ctxs[j].detect_enqueue(img, W, H, channels=3)
                    pending[j] = True
                for j in range(n_ctx):
                    if pending[j]:
                        ctxs[j].result_copy(outs[j])

            one_step_rgb()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps_rgb = 3
            for _ in range(reps_rgb):
                one_step_rgb()
            dt = time.perf_counter() - t0
            e2e_rgb = {"value": n_rgb * reps_rgb / dt, "unit": "images/s", "images": n_rgb, "channels": 3,
                       "h2d_bytes_per_image": 3 * W * H,
                       "note": "pinned RGB u8 in, keypoint records out; gray conversion fused into the input kernel"}
            del h_rgb
        except Exception as ex:
            e2e_rgb = {"error": repr(ex)}

    # ---- per-stage device times + roofline of the dominant (pyramid) kernel: one context, one
    # stream, events around every stage, same images ----
    roof = stages = stage_launches = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if rank == 0:
        c0 = ctxs[0]
        c0.set_profiling(True)
        acc = None
        marks_acc = None
        reps = min(len(d_imgs), 16)
        for k in range(reps + 2):
            c0.detect_enqueue(d_imgs[k % len(d_imgs)], W, H)
            ms, nl = c0.profile()
            marks = c0.profile_marks()
            if k >= 2:
                acc = ms if acc is None else {s: acc[s] + ms[s] for s in ms}
                marks_acc = [t for _, t in marks] if marks_acc is None else [a + t for a, (_, t) in zip(marks_acc, marks)]
        c0.set_profiling(False)
        stages = {s: acc[s] / reps for s in acc}
        stage_launches = nl
        pyr_bytes = pyramid_algorithmic_bytes(W, H)
        hbm_peak = peaks.get("hbm_gbs")
        peak_src = "MEASURED_PEAKS.json hbm_gbs (sustained copy)" if hbm_peak else "fallback 6650 GB/s (B200_PROFILING.md)"
        hbm_peak = hbm_peak or 6650.0
        achieved = pyr_bytes / (stages["pyramid"] * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("pyramid_dram_bytes_per_image")
        except Exception:
            pass
        total_ms = sum(stages.values())
        # per-kernel view of octave 0 (the two largest launches): live event times of the two fused kernels;
        # algorithmic bytes per octave pixel: first kernel reads G0 (4), writes G1..G3, D0..D2 (24) and the next
        # base (1); second kernel reads G3 (4), writes D3, D4 (8)
        per_kernel = None
        pyr_marks = [i for i, (st, _) in enumerate(marks) if st == "pyramid"]
        fused = (len(pyr_marks) >= 2 and pyr_marks[1] == pyr_marks[0] + 1 and pyr_marks[1] + 1 < len(marks)
                 and marks[pyr_marks[1] + 1][0] == "extrema")
        tail = None
        if fused:   # fused path: octave 0 kernel by kernel
            bw, bh = (2 * W, 2 * H)
            px0 = bw * bh
            per_kernel = []
            for name, idx, bpp in (("octave 0: G0 -> G1..G3, D0..D2, next base", pyr_marks[0], 29.0),
                                   ("octave 0: G3 -> (G4, G5 on chip) -> D3, D4", pyr_marks[1], 12.0)):
                t = marks_acc[idx] / reps
                per_kernel.append({"kernel": name, "ms": t, "algorithmic_bytes": bpp * px0,
                                   "achieved": bpp * px0 / (t * 1e-3) / 1e9, "frac": bpp * px0 / (t * 1e-3) / 1e9 / hbm_peak})
            ex0 = pyr_marks[1] + 1   # the mark after octave 0's second cascade kernel is octave 0's extrema scan
            if ex0 < len(marks) and marks[ex0][0] == "extrema":
                t = marks_acc[ex0] / reps
                per_kernel.append({"kernel": "octave 0: 3x3x3 extrema scan (k_extrema4), reads D0..D4", "ms": t,
                                   "algorithmic_bytes": 20.0 * px0, "achieved": 20.0 * px0 / (t * 1e-3) / 1e9,
                                   "frac": 20.0 * px0 / (t * 1e-3) / 1e9 / hbm_peak})
            octs = c0.stats()["octaves"]
            if nl["pyramid"] % 2 == 1 and pyr_marks[-1] + 1 < len(marks):   # ... + one launch for the small octaves
                n_tail = octs - (nl["pyramid"] - 1) // 2
                tail = {"kernel": "k_tail: both cascade kernels of the last %d octaves in one ticket-ordered launch" % n_tail,
                        "ms": marks_acc[pyr_marks[-1]] / reps, "octaves": n_tail,
                        "extrema_ms": marks_acc[pyr_marks[-1] + 1] / reps if marks[pyr_marks[-1] + 1][0] == "extrema" else None}
        roof = {"bound": "hbm", "kernel": "pyramid: fused cascade G0->G1..G3,D0..D2,next base + G3->D3,D4 (k_stream on octaves >= 2 Mpx, k_cascade below, k_tail for the octaves of <= 1 tile per SM), all octaves of one image", "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "algorithmic_bytes": pyr_bytes, "ms": stages["pyramid"], "launches": nl["pyramid"],
                "peak_source": peak_src, "per_kernel": per_kernel, "tail": tail,
                "whole_detect": {"algorithmic_bytes": HBM_BYTES_PER_INPUT_PIXEL * W * H, "ms": total_ms,
                                 "frac": HBM_BYTES_PER_INPUT_PIXEL * W * H / (total_ms * 1e-3) / 1e9 / hbm_peak}}

    # ---- the matcher (the only tensor-core stage): one 20k x 20k pair of config 5, device-resident ----
    match = None
    if rank == 0:
        nm = 20000
        da, db = synth_desc_gpu(nm, 1234, dev), synth_desc_gpu(nm, 1235, dev)
        mi = torch.empty(nm, dtype=torch.int32, device=dev); m1 = torch.empty_like(mi); m2 = torch.empty_like(mi)
        torch.cuda.synchronize()
        c0 = ctxs[0]
        for _ in range(5):
            c0.match_enqueue(da, nm, db, nm, mi, m1, m2)
        c0.sync()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for _ in range(reps):
            c0.match_enqueue(da, nm, db, nm, mi, m1, m2)
        e1.record(streams[0])
        c0.sync()
        us = e0.elapsed_time(e1) * 1e3 / reps
        tf = 2.0 * nm * nm * 128 / (us * 1e-6) / 1e12
        tpeak = peaks.get("bf16_tflops") or 1590.0
        match = {"workload": "20000 x 20000 x 128 u8 descriptors, top-2 + norms fused (tcgen05 kind::i8)",
                 "us_per_pair": us, "tflops_equivalent": tf, "peak_bf16_tflops": tpeak, "frac_of_bf16_peak": tf / tpeak,
                 "frac_of_i8_peak": tf / (2.0 * tpeak),
                 "i8_peak_note": "kind::i8 runs at twice the bf16 rate; 2 x the measured bf16 peak is used (no measured i8 peak on file)",
                 "path": "tcgen05" if c0.match_path(nm, nm) else "simt", "data_in_l2": True,
                 "note": "2*N1*N2*128 / t; both descriptor sets (5 MB) are L2-resident by construction of the problem"}
        # the same kernel with operands that do not fit L2: 2 x 77 MB of descriptors, every 128-row block of A streams
        # all of B (600 000 x 128 B) through the TMA ring
        try:
            nl = 600000
            la, lb = synth_desc_gpu(nl, 4321, dev), synth_desc_gpu(nl, 4322, dev)
            li = torch.empty(nl, dtype=torch.int32, device=dev); l1 = torch.empty_like(li); l2 = torch.empty_like(li)
            torch.cuda.synchronize()
            c0.match_enqueue(la, nl, lb, nl, li, l1, l2)
            c0.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(streams[0])
            c0.match_enqueue(la, nl, lb, nl, li, l1, l2)
            e1.record(streams[0])
            c0.sync()
            msl = e0.elapsed_time(e1)
            tfl = 2.0 * nl * nl * 128 / (msl * 1e-3) / 1e12
            match["hbm_resident"] = {"workload": f"{nl} x {nl} descriptors (2 x {nl * 128 / 1e6:.0f} MB, larger than the 126 MB L2)",
                                     "ms": msl, "tflops_equivalent": tfl, "frac_of_bf16_peak": tfl / tpeak,
                                     "frac_of_i8_peak": tfl / (2.0 * tpeak)}
            del la, lb, li, l1, l2
        except Exception as ex:
            match["hbm_resident"] = {"error": repr(ex)}

    # ---- single-image latency: one context, one image at a time, device-resident input, graph replay ----
    latency = None
    if rank == 0:
        c0 = ctxs[0]
        for k in range(3):
            c0.detect_enqueue(d_imgs[k % len(d_imgs)], W, H)
            c0.detect_finish()
        reps = min(len(d_imgs), 16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for k in range(reps):
            c0.detect_enqueue(d_imgs[k % len(d_imgs)], W, H)
        e1.record(streams[0])
        c0.detect_finish()
        back_to_back = e0.elapsed_time(e1) / reps
        t0 = time.perf_counter()
        for k in range(reps):
            c0.detect_enqueue(d_imgs[k % len(d_imgs)], W, H)
            c0.detect_finish()
        latency = {"ms_per_image_one_stream": back_to_back, "ms_enqueue_to_finish_host": 1e3 * (time.perf_counter() - t0) / reps,
                   "how": "one context; CUDA events around 16 back-to-back detect calls / host clock around enqueue + finish",
                   "graph": bool(int(os.environ.get("SIFT_B200_GRAPH", "1")))}

    # ---- config 5 in miniature on every run (and at every N of the scaling run): all-pairs matching of 32 x 8000
    # descriptors through sift_b200_collection_match; at N > 1 the descriptors cross NVLink in one ncclAllGather
    # issued by the library, and the all-reduced digest must equal the digest of one context doing it all ----
    collection = None
    try:
        attach_comm(ctxs[0], rank, world)
        c_sets, c_per = 32, 8000
        ms_c, n_pairs_c, matches_c, digest_c, solo_c = collection_pass(ctxs[0], streams[0], c_sets, c_per, rank, world,
                                                                       dev, steps=3, warmup=1, check_solo=True)
        ms_c = reduce_max(ms_c)
        if rank == 0:
            flop = 2.0 * (c_sets * (c_sets - 1) // 2) * c_per * c_per * 128
            collection = {"workload": f"{c_sets} sets x {c_per} descriptors, all {c_sets * (c_sets - 1) // 2} pairs, through "
                                      "sift_b200_collection_match (NCCL all-gather inside the library at N > 1)",
                          "ms": ms_c, "tflops_equivalent": flop / (ms_c * 1e-3) / 1e12, "pairs_on_rank0": n_pairs_c,
                          "matches": matches_c, "digest": f"{digest_c:016x}",
                          "one_gpu_digest": f"{solo_c[1]:016x}", "equals_one_gpu_run": (matches_c, digest_c) == tuple(solo_c)}
    except Exception as ex:
        collection = {"error": repr(ex)}

    cpu = parity = config1 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, parity = cpu_baseline_sample(h_imgs[0].numpy(), local)
        except Exception as ex:  # the oracle is only a reported baseline; never fatal
            cpu = {"error": repr(ex)}
        try:
            config1 = config1_pair(local)
        except Exception as ex:
            config1 = {"error": repr(ex)}
        try:
            if match is not None:
                match["cpu_baseline"] = cpu_match_baseline()
        except Exception as ex:
            match["cpu_baseline"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic {W}x{H} batch of {batch} images, detect+describe, sharded by image",
                       "global_batch": batch, "images_per_rank": len(mine), "contexts_per_gpu": n_ctx,
                       "generator": "D (sum of Gaussian-filtered noise, sigma 2..32), seed 1234+i",
                       "keypoints_per_image": kp_mean, "octaves": ctxs[0].stats()["octaves"],
                       "l2": "inputs larger than L2 (each image's pyramid is ~2 GB; 126 MB L2)",
                       "parallelism": f"dp{world} by image, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_step,
                    "d2h_bytes_per_step": d2h_step, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "stages_ms": stages, "stage_launches": stage_launches, "match": match, "latency": latency,
            "parity": parity, "config1": config1, "collection": collection, "e2e_rgb": e2e_rgb,
        }
        emit(line)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


def synth_desc_gpu(n, seed, dev):
    """Config-5 style descriptors on the GPU: |N(0,1)| through the reference's normalise -> clamp 0.2 ->
    renormalise -> floor(512 x) -> min 255 (sift.cpp:582-602)."""
    import torch
    g = torch.Generator(device=dev); g.manual_seed(seed)
    hh = torch.randn((n, 128), generator=g, device=dev).abs()
    hh = hh / hh.norm(dim=1, keepdim=True)
    hh = hh.clamp(max=0.2)
    hh = hh / hh.norm(dim=1, keepdim=True)
    return torch.floor(512.0 * hh).clamp(max=255).to(torch.uint8).contiguous()


def attach_comm(ctx, rank, world):
    """NCCL communicator inside libsift_b200.so: rank 0's unique id travels over torch.distributed (plumbing)."""
    import torch.distributed as dist
    import sift_project_b200 as S
    if world == 1:
        ctx.comm_attach(None, 1, 0)
        return
    uid = [S.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_attach(uid[0], world, rank)


def collection_pass(ctx, stream, n_sets, per, rank, world, dev, steps=2, warmup=1, check_solo=False):
    """All-pairs matching of `n_sets` x `per` descriptors through the C ABI (sift_b200_collection_match: counts
    all-reduce + ONE ncclAllGather on the side stream overlapped with the local pairs + every owned pair's top-2 on
    tcgen05).  Returns (ms per pass -- CUDA events on the context's stream --, matches, digest, solo digest)."""
    import torch
    import sift_project_b200 as S
    owned = [synth_desc_gpu(per, 1234 + i, dev) for i in range(n_sets) if i % world == rank]
    torch.cuda.synchronize()
    for _ in range(warmup):
        ctx.collection_match(n_sets, owned)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        n_pairs = ctx.collection_match(n_sets, owned)
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / steps
    matches, digest = ctx.collection_digest(0.75, all_ranks=True)
    solo = None
    if check_solo and rank == 0:      # the whole collection on ONE context: same digest or the exchange is wrong
        everything = [synth_desc_gpu(per, 1234 + i, dev) for i in range(n_sets)]
        with S.SiftContext(64, 64, device=dev.index) as one:
            one.collection_match(n_sets, everything)
            solo = one.collection_digest(0.75, all_ranks=False)
    return ms, n_pairs, matches, digest, solo


def run_collection(args):
    """BASELINE.json config 5: `--sets` descriptor sets x `--per-set` descriptors, all-pairs matching partitioned
    over the ranks by image pair, after ONE NCCL all-gather of the u8 descriptor blocks -- all inside the C ABI.
    Timed region (CUDA events, max over ranks): all-gather + every pair's top-2 search."""
    import torch
    import torch.distributed as dist
    import sift_project_b200 as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n_sets, per = args.sets, args.per_set
    ctx = S.SiftContext(64, 64, device=local)
    attach_comm(ctx, rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    if world > 1:
        dist.barrier(device_ids=[local])
    ms_local, n_pairs, matches, digest, _ = collection_pass(ctx, stream, n_sets, per, rank, world, dev,
                                                            steps=args.steps, warmup=min(args.warmup, 1))
    ms = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        total_pairs = n_sets * (n_sets - 1) // 2
        flop = 2.0 * total_pairs * per * per * 128
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tpeak = (peaks.get("bf16_tflops_sustained") or 1400.0) * world
        tf = flop / (float(ms.item()) * 1e-3) / 1e12
        emit(({
            "metric": "collection all-pairs match (top-2 + ratio test)", "value": tf, "unit": "TFLOP/s-equivalent",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms.item()),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{n_sets} sets x {per} descriptors, {total_pairs} unordered pairs, "
                                   f"ncclAllGather of {n_sets * per * 128 / 1e9:.2f} GB inside libsift_b200.so + "
                                   f"pair-partitioned matching", "pairs_on_rank0": n_pairs},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                         "traffic": None, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained x n_gpus"},
            "matches": matches, "digest": f"{digest:016x}"}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL's version banner,
    torchrun notices) are routed to stderr; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=256, help="global batch (config: 256)")
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--contexts", type=int, default=4, help="contexts (streams) in flight per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="detect", choices=["detect", "collection"],
                    help="detect = the headline metric; collection = config 5 (all-pairs matching over N GPUs)")
    ap.add_argument("--sets", type=int, default=512)
    ap.add_argument("--per-set", type=int, default=20000)
    args = ap.parse_args()
    if args.workload == "collection":
        run_collection(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
