mkdir -p gpurun_out
B="timeout 300 python bench.py --images 64 --steps 3 --warmup 3 --no-cpu-baseline"
for g in 0 1 2 3; do
SIFT_B200_STREAM_A=$g $B > gpurun_out/b_r2l_a$g.json 2> gpurun_out/b_r2l_a$g.err
done
for f in a0 a1 a2 a3; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2l_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), 'pyramid %.3f'%d['stages_ms']['pyramid'], 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'], [round(p['ms'],4) for p in d['roofline']['per_kernel']], 'kp', d['config']['keypoints_per_image'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2l_'+f+'.err').read()[-600:])
PY
done
