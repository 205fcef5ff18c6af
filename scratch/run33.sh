mkdir -p gpurun_out
SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_old.log 2>&1
timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_new.log 2>&1
tail -3 gpurun_out/ab_old.log; tail -3 gpurun_out/ab_new.log
cmp gpurun_out/ab_old.log gpurun_out/ab_new.log && echo "AB IDENTICAL"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),{k:round(v,4) for k,v in d["stages_ms"].items()})
PY
