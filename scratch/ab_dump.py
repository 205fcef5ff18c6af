import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sift_project_b200 as S
from oracle import oracle as O
tag = sys.argv[1]
with S.SiftContext(1024, 768) as c:
    img = O.synth_image(768, 1024, seed=9)
    k = c.detect(img)
    np.save(f"gpurun_out/ab_{tag}_768.npy", k)
    print(tag, len(k))
