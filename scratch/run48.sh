mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_final_default.json 2> gpurun_out/b_final_default.err
timeout 300 python bench.py --width 1920 --height 1080 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_final_1080p.json 2> gpurun_out/b_final_1080p.err
for f in default 1080p; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_final_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms'].items()}, 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'], 'roof %.3f'%d['roofline']['frac'], d['clocks']['reasons'], d.get('parity',{}) and d['parity'].get('desc_max'))
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_final_'+f+'.err').read()[-600:])
PY
done
