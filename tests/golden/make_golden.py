"""Regenerates the golden vectors in this directory from the REAL reference
(oracle/_ref/libsift_ref.so, compiled from /root/reference/src by oracle/Makefile).

Run in the build container only (needs /root/reference):
    make -C oracle && python tests/golden/make_golden.py

Files written
  image1.png, image2.png   the reference's stitching/image{1,2}.jpg decoded by the reference's own
                           vendored stb_image (image_io.cpp:20-35) and re-saved losslessly -- a
                           different JPEG decoder gives different pixels, so the decode is pinned.
  config1.npz              stage counts, final keypoints (168-byte records) of both images and the
                           image1->image2 match list (SURVEY.md section 4 known answers).
  synth_256x192.npz        generator-D image (seed 42): stage counts, extrema, raw / oriented /
                           final keypoints, per-layer checksums of the FP64 Gaussian and DoG planes.
  match_kat.npz            matcher known answers on synthetic descriptors incl. ties, |B| = 0, 1.
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_IMG = "/root/reference/stitching/{}.jpg"


def strip_desc(k):
    """Stages before compute_descriptors carry uninitialised desc bytes in the reference."""
    k = k.copy()
    k["desc"] = 0
    return k


def stage_dump(run):
    return dict(
        octaves=run.octaves,
        sigmas=run.sigmas(),
        extrema=run.extrema().astype(np.int32),
        raw=strip_desc(run.keypoints(0)),
        oriented=strip_desc(run.keypoints(1)),
        final=run.keypoints(2),
    )


def main():
    lib = O.ref()
    out = {}
    kps = []
    for name in ("image1", "image2"):
        px = O.ref_load_image(REF_IMG.format(name))
        u8 = px.astype(np.uint8)
        assert np.array_equal(u8.astype(np.float64), px)
        Image.fromarray(u8).save(os.path.join(HERE, name + ".png"), optimize=True)
        run = O.Run(lib, px, keep_pyramid=False)
        d = stage_dump(run)
        for k, v in d.items():
            out[f"{name}_{k}"] = v
        kps.append(d["final"])
        print(name, len(d["extrema"]), len(d["raw"]), len(d["oriented"]), len(d["final"]))
    ia, ib, dist = O.match(lib, kps[0]["desc"], kps[1]["desc"])
    out["match_ia"], out["match_ib"], out["match_dist"] = ia, ib, dist
    print("matches", len(ia))
    np.savez_compressed(os.path.join(HERE, "config1.npz"), **out)

    img = O.synth_image(192, 256, seed=42)
    run = O.Run(lib, img, keep_pyramid=True)
    d = stage_dump(run)
    d["image"] = img
    sums = []
    for o in range(run.octaves):
        for l in range(6):
            g = run.gaussian(o, l)
            sums.append((o, l, 0, g.sum(), np.abs(g).max(), g[g.shape[0] // 2, g.shape[1] // 2]))
        for l in range(5):
            g = run.dog(o, l)
            sums.append((o, l, 1, g.sum(), np.abs(g).max(), g[g.shape[0] // 2, g.shape[1] // 2]))
    d["plane_checks"] = np.array(sums)
    # one full-resolution plane pair for element-wise checks of the port
    d["g_o1_l3"] = run.gaussian(1, 3)
    d["dog_o1_l2"] = run.dog(1, 2)
    np.savez_compressed(os.path.join(HERE, "synth_256x192.npz"), **d)
    print("synth", len(d["extrema"]), len(d["raw"]), len(d["oriented"]), len(d["final"]))

    # matcher known answers: ties (duplicate rows in B), duplicates of the best, tiny sets
    a = O.synth_descriptors(300, seed=1)
    b = O.synth_descriptors(257, seed=2)
    b[100] = b[7]          # exact duplicate rows -> best == second, no match, lowest j wins
    b[200] = a[5]          # an exact hit for a[5]
    b[201] = a[5]          # ... twice: distance 0 tie -> 0 < 0.75*0 is false
    a[11] = b[33]          # unique exact hit
    rng = np.random.default_rng(3)
    for t in range(60):    # near-duplicates: most pass the ratio test
        b[120 + t] = np.clip(a[2 * t + 20].astype(np.int32) + rng.integers(-6, 7, 128), 0, 255)
    m = {}
    for tag, (x, y) in {"full": (a, b), "b1": (a[:20], b[:1]), "b0": (a[:20], b[:0]),
                        "a0": (a[:0], b), "b2": (a[:50], b[:2])}.items():
        ia, ib, dist = O.match(lib, x, y)
        m[f"{tag}_ia"], m[f"{tag}_ib"], m[f"{tag}_dist"] = ia, ib, dist
        print("kat", tag, len(ia))
    m["a"], m["b"] = a, b
    np.savez_compressed(os.path.join(HERE, "match_kat.npz"), **m)


if __name__ == "__main__":
    main()
