"""The PNGs the REFERENCE writes for config 1 -- keypoints.png (sift.cpp:765-768, overwritten by every detect call,
so it shows image 2) and matches.png (sift.cpp:850-876) -- stored as the overlay on top of the input pixels
(ref_drawings.npz: flat pixel indices + RGB where the drawing differs from image2.png / [image1 | image2]).

Run in the build container only: the as-shipped reference needs ~6.5 minutes for this pair (it deep-copies whole
octaves per extremum, sift.cpp:346).
    make -C oracle && python tests/golden/make_golden_drawings.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "sift")
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call([exe, os.path.join(HERE, "image1.png"), os.path.join(HERE, "image2.png")], cwd=d,
                              stdout=subprocess.DEVNULL)
        kp = np.asarray(Image.open(os.path.join(d, "keypoints.png")))
        mt = np.asarray(Image.open(os.path.join(d, "matches.png")))
    i1 = np.asarray(Image.open(os.path.join(HERE, "image1.png")))
    i2 = np.asarray(Image.open(os.path.join(HERE, "image2.png")))
    ik = np.flatnonzero(np.any(kp != i2, -1)).astype(np.int32)
    im = np.flatnonzero(np.any(mt != np.concatenate([i1, i2], 1), -1)).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_drawings.npz"), kp_shape=np.array(kp.shape), kp_idx=ik,
                        kp_rgb=kp.reshape(-1, 3)[ik], mt_shape=np.array(mt.shape), mt_idx=im, mt_rgb=mt.reshape(-1, 3)[im])
    print(len(ik), len(im), "overlay pixels")


if __name__ == "__main__":
    sys.exit(main())
