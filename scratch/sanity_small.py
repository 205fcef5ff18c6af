"""Small end-to-end run for compute-sanitizer (development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sift_project_b200 as S
from oracle import oracle as O
img = O.synth_image(150, 210, seed=3)
with S.SiftContext(256, 192) as c:
    for kw in (dict(), dict(double_image_size=False), dict(intervals=4)):
        k = c.detect(img, **kw)
        print(kw, len(k), c.stats())
    rgb = np.stack([img, img, img], -1)
    print("rgb", len(c.detect(rgb)))
    a, b = O.synth_descriptors(300, 1), O.synth_descriptors(700, 2)
    os.environ["SIFT_B200_MATCH"] = "tc"
    print("tc", len(c.match(a, b)[0]))
    os.environ["SIFT_B200_MATCH"] = "simt"
    print("simt", len(c.match(a, b)[0]))
