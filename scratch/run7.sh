mkdir -p gpurun_out
export SIFT_B200_STREAM_TMA=1
SIFT_B200_PYRAMID_MODE=3 timeout 120 python - > gpurun_out/tma_small.txt 2>&1 <<'PY'
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
rc=$?
echo "small exit $rc" >> gpurun_out/tma_small.txt
tail -3 gpurun_out/tma_small.txt
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --maxfail=3 -p no:cacheprovider -k "fused or stream or random_shapes or 4k_set or stage_counts or graph or rgb or reproducible" > gpurun_out/pytest_r2f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2f.log
tail -6 gpurun_out/pytest_r2f.log
B="timeout 300 python bench.py --images 32 --steps 3 --warmup 3 --no-cpu-baseline"
SIFT_B200_STREAM_TMA=1 $B > gpurun_out/b_r2f_tma1.json 2> gpurun_out/b_r2f_tma1.err
SIFT_B200_STREAM_TMA=0 $B > gpurun_out/b_r2f_tma0.json 2> gpurun_out/b_r2f_tma0.err
SIFT_B200_STREAM_TMA=1 $B --contexts 6 > gpurun_out/b_r2f_tma1_c6.json 2> gpurun_out/b_r2f_tma1_c6.err
SIFT_B200_STREAM_TMA=0 $B --contexts 6 > gpurun_out/b_r2f_tma0_c6.json 2> gpurun_out/b_r2f_tma0_c6.err
SIFT_B200_STREAM_TMA=0 $B --contexts 3 > gpurun_out/b_r2f_tma0_c3.json 2> gpurun_out/b_r2f_tma0_c3.err
for f in tma1 tma0 tma1_c6 tma0_c6 tma0_c3; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2f_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms'].items()}, 'lat %.3f'%d['latency']['ms_per_image_one_stream'], [round(p['ms'],4) for p in d['roofline']['per_kernel']])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2f_'+f+'.err').read()[-600:])
PY
done
