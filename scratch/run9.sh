mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -rf --maxfail=5 -p no:cacheprovider -s -k "batch or graph or fused_octave or drop_in or reproducible or streaming" > gpurun_out/pytest_r2h.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2h.log
grep -E "passed|failed|batch:|differ" gpurun_out/pytest_r2h.log | tail -8
B="timeout 300 python bench.py --images 64 --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/b_r2h_default.json 2> gpurun_out/b_r2h_default.err
SIFT_B200_FORK_MIN_PX=0 $B > gpurun_out/b_r2h_fork0.json 2> gpurun_out/b_r2h_fork0.err
SIFT_B200_FORK_MIN_PX=4000000 $B > gpurun_out/b_r2h_fork4m.json 2> gpurun_out/b_r2h_fork4m.err
SIFT_B200_FORK_MIN_PX=300000 $B > gpurun_out/b_r2h_fork300k.json 2> gpurun_out/b_r2h_fork300k.err
for f in default fork0 fork4m fork300k; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/b_r2h_'+f+'.json').read().strip().splitlines()[-1])
    print(f, 'value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), 'sum %.3f'%sum(d['stages_ms'].values()), 'lat %.3f'%d['latency']['ms_per_image_one_stream'])
except Exception as e:
    print(f,'ERR',e, open('gpurun_out/b_r2h_'+f+'.err').read()[-600:])
PY
done
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -s 41 -c 41 --csv --log-file gpurun_out/launches_r2h.csv python scratch/one_detect.py 3 > gpurun_out/ncu_r2h_launches.log 2>&1
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_describe -s 1 -c 1 -o gpurun_out/prof_r2h_describe python scratch/one_detect.py 3 > gpurun_out/ncu_r2h.log 2>&1
tail -2 gpurun_out/ncu_r2h.log
