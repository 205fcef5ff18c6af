mkdir -p gpurun_out
B="timeout 300 python bench.py --images 32 --steps 3 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/b_r2s_$name.json 2> gpurun_out/b_r2s_$name.err; python - $name <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2s_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f'%d['value'], 'pyramid %.3f extrema %.3f'%(d['stages_ms']['pyramid'], d['stages_ms']['extrema']), 'lat %.3f'%d['latency']['ms_per_image_one_stream'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2s_'+f+'.err').read()[-300:])
PY
}
run base X=1
run bn3m SIFT_B200_STREAM_B_NARROW_PX=3000000
run bn10m SIFT_B200_STREAM_B_NARROW_PX=10000000
run exr8 SIFT_B200_EX_ROWS_MID=8
run exr32 SIFT_B200_EX_ROWS_MID=32
run smin1m SIFT_B200_STREAM_MIN_PX=1000000
run smin400k SIFT_B200_STREAM_MIN_PX=400000
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "match" > gpurun_out/pytest_r2s.log 2>&1; tail -2 gpurun_out/pytest_r2s.log
