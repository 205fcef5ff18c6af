mkdir -p gpurun_out
timeout 60 ./scratch/tma_probe > gpurun_out/tma_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/tma_probe.txt
cat gpurun_out/tma_probe.txt
SIFT_B200_STREAM_TMA=1 SIFT_B200_PYRAMID_MODE=3 timeout 120 python - > gpurun_out/tma_small.txt 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import sift_project_b200 as S
from oracle import oracle as O
img = O.synth_image(192, 256, seed=42)
with S.SiftContext(256, 192) as c:
    c.launch_plan(use_graph=0)
    k = c.detect(img)
    print("tma mode 3 keypoints", len(k))
PY
echo "small exit $?" >> gpurun_out/tma_small.txt
cat gpurun_out/tma_small.txt | tail -5
