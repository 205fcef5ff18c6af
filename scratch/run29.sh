mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_r2x_default.json 2> gpurun_out/b_r2x_default.err
timeout 300 python bench.py --width 7680 --height 4320 --images 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2x_8k.json 2> gpurun_out/b_r2x_8k.err
timeout 300 python bench.py --width 1920 --height 1080 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2x_1080p.json 2> gpurun_out/b_r2x_1080p.err
for f in default 8k 1080p; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2x_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'], 'roof %.3f'%d['roofline']['frac'], [(round(p['ms'],4),round(p['frac'],3)) for p in (d['roofline']['per_kernel'] or [])], d['roofline'].get('tail'))
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2x_'+f+'.err').read()[-600:])
PY
done
