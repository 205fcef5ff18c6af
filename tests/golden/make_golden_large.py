"""Golden keypoints of the REAL reference (oracle/_ref/libsift_ref.so, copy-free build, bit-identical to the
as-shipped one) at BASELINE.json's full sizes, so that the GPU tests can compare SETS -- not counts -- at 4K and
8K without spending minutes of GPU-box time in the CPU reference (4K: ~2 min, 4 GB; 8K: ~8 min, 17 GB).

Run in the build container only (needs /root/reference and the RAM):
    make -C oracle && python tests/golden/make_golden_large.py [4k] [8k]

Files written
  synth_3840x2160_seed1234.npz   config 3's image 0 (generator D, seed 1234): stage counts and EVERY final
                                 keypoint: x, y, size, pori (float64), octave, layer, and the 128-byte descriptor.
  synth_7680x4320_seed1234.npz   config 4: stage counts, every final keypoint's x, y, size, pori, octave, layer,
                                 and the descriptors of every 8th keypoint (the full set would be 22 MB).
The images themselves are regenerated from the seed by oracle.synth_image (numpy + scipy, deterministic).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(h, w, seed, desc_stride):
    img = O.synth_image(h, w, seed=seed)
    t = time.time()
    run = O.Run(O.ref(), img, keep_pyramid=False)
    k = run.keypoints(2)
    counts = np.array([run.octaves, len(run.extrema()), len(run.keypoints(0)), len(run.keypoints(1)), len(k)])
    # (counts[0] is 0: the octave count is read from the pyramid, which keep_pyramid=False has already released)
    print(f"{w}x{h}: octaves/extrema/raw/oriented/final = {counts.tolist()}  ({time.time() - t:.0f} s)")
    out = dict(counts=counts, x=k["x"], y=k["y"], size=k["size"], pori=k["pori"],
               octave=k["octave"].astype(np.int8), layer=k["layer"].astype(np.int8),
               desc=np.ascontiguousarray(k["desc"][::desc_stride]), desc_stride=np.array(desc_stride),
               image_crc=np.array(int(np.uint64(img.astype(np.uint64).sum()))))
    np.savez_compressed(os.path.join(HERE, f"synth_{w}x{h}_seed{seed}.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["4k", "8k"]
    if "4k" in which:
        dump(2160, 3840, 1234, 1)
    if "8k" in which:
        dump(4320, 7680, 1234, 8)
