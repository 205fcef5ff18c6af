mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --maxfail=10 -p no:cacheprovider -k "fused or stream or random_shapes or 4k_set or stage_counts or graph or rgb" > gpurun_out/pytest_r2e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2e.log
tail -6 gpurun_out/pytest_r2e.log
B="python bench.py --images 32 --steps 3 --warmup 3 --no-cpu-baseline"
SIFT_B200_STREAM_TMA=1 $B > gpurun_out/b_r2e_tma1.json 2> gpurun_out/b_r2e_tma1.err
SIFT_B200_STREAM_TMA=0 $B > gpurun_out/b_r2e_tma0.json 2> gpurun_out/b_r2e_tma0.err
SIFT_B200_STREAM_TMA=1 $B > gpurun_out/b_r2e_tma1b.json 2> gpurun_out/b_r2e_tma1b.err
SIFT_B200_STREAM_TMA=0 $B > gpurun_out/b_r2e_tma0b.json 2> gpurun_out/b_r2e_tma0b.err
for f in tma1 tma0 tma1b tma0b; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2e_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms'].items()}, 'lat %.3f'%d['latency']['ms_per_image_one_stream'], [round(p['ms'],4) for p in d['roofline']['per_kernel']])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2e_'+f+'.err').read()[-600:])
PY
done
