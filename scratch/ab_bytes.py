"""Digest of detect results on a set of images: run with SIFT_B200_LIB pointing at two builds and compare."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sift_project_b200 as S
from oracle import oracle as O
h = hashlib.sha256()
shapes = [(2160, 3840, 1234), (1080, 1920, 7), (768, 1024, 9), (301, 517, 3), (97, 131, 5)]
with S.SiftContext(3840, 2160) as c:
    for hh, ww, seed in shapes:
        img = O.synth_image(hh, ww, seed=seed)
        for kw in ({}, {"double_image_size": False}):
            k = c.detect(img, **kw)
            h.update(k.tobytes())
            print(hh, ww, kw, len(k), hashlib.sha256(k.tobytes()).hexdigest()[:16])
    rgb = np.random.default_rng(1).integers(0, 256, (480, 640, 3), dtype=np.uint8)
    k = c.detect(rgb); h.update(k.tobytes())
    f = (O.synth_image(240, 320, seed=41).astype(np.float32) * np.float32(0.01) - np.float32(1.0))
    k = c.detect(f, contrast_threshold=0.0004); h.update(k.tobytes())
    print("f32", len(k))
print("DIGEST", h.hexdigest())
