mkdir -p gpurun_out
B="timeout 300 python bench.py --images 16 --steps 2 --warmup 3 --no-cpu-baseline"
for v in c4p1 c4p4 c4p8 c4p9 c4p12 c4p20 c2p4 c2p16 c2p17 c4p8b6; do
SIFT_B200_LIB=$PWD/scratch/variants/libsift_b200_$v.so $B > gpurun_out/b_r2o_$v.json 2> gpurun_out/b_r2o_$v.err
python - $v <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2o_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f'%d['value'], 'describe %.3f'%d['stages_ms']['describe'], 'lat %.3f'%d['latency']['ms_per_image_one_stream'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2o_'+f+'.err').read()[-300:])
PY
done
