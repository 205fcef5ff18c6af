"""One 4K detect, a few times (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import sift_project_b200 as S
W, H = 3840, 2160
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
img = bench.synth_image_gpu(H, W, 1234, dev)
torch.cuda.synchronize()
with S.SiftContext(W, H) as c:
    c.launch_plan(use_graph=0)
    for _ in range(n):
        c.detect_enqueue(img, W, H)
        k = c.detect_finish()
    print("keypoints", k, c.stats())
