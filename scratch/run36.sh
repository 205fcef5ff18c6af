mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_r2w2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2w2.log
tail -4 gpurun_out/pytest_r2w2.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_r2z_default.json 2> gpurun_out/b_r2z_default.err
timeout 300 python bench.py --width 7680 --height 4320 --images 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2z_8k.json 2> gpurun_out/b_r2z_8k.err
timeout 300 python bench.py --width 1920 --height 1080 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2z_1080p.json 2> gpurun_out/b_r2z_1080p.err
for f in default 8k 1080p; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2z_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms'].items()}, 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'], 'roof %.3f'%d['roofline']['frac'], [(round(p['ms'],4),round(p['frac'],3)) for p in (d['roofline']['per_kernel'] or [])], d['clocks'])
    if d.get('parity'): print('   parity', d['parity'])
    if d.get('e2e_rgb'): print('   e2e_rgb', d['e2e_rgb'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2z_'+f+'.err').read()[-600:])
PY
done
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -s 25 -c 25 --csv --log-file gpurun_out/launches_r2z.csv python scratch/one_detect.py 3 > gpurun_out/ncu_r2z_launches.log 2>&1
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_orient|k_refine|k_extrema4|k_stream|k_input_u8" -s 17 -c 8 -o gpurun_out/prof_r2z_top python scratch/one_detect.py 3 > gpurun_out/ncu_r2z_d.log 2>&1
tail -1 gpurun_out/ncu_r2z_d.log
