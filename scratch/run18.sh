mkdir -p gpurun_out
B="timeout 300 python bench.py --images 16 --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/b_r2p_base.json 2> gpurun_out/b_r2p_base.err
for v in o16s66 o16s38 o32s37 o8s37; do
SIFT_B200_LIB=$PWD/scratch/variants/libsift_b200_$v.so $B > gpurun_out/b_r2p_$v.json 2> gpurun_out/b_r2p_$v.err
done
for v in base o16s66 o16s38 o32s37 o8s37; do
python - $v <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2p_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f'%d['value'], 'describe %.3f orient %.3f'%(d['stages_ms']['describe'], d['stages_ms']['orient']), 'lat %.3f'%d['latency']['ms_per_image_one_stream'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2p_'+f+'.err').read()[-300:])
PY
done
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_orient|k_refine" -s 3 -c 3 -o gpurun_out/prof_r2p python scratch/one_detect.py 3 > gpurun_out/ncu_r2p.log 2>&1
tail -2 gpurun_out/ncu_r2p.log
