mkdir -p gpurun_out
SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_old.log 2>&1
timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_new.log 2>&1
cmp gpurun_out/ab_old.log gpurun_out/ab_new.log && echo "AB IDENTICAL"
timeout 1500 python -m pytest tests -m gpu -q -rf --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),{k:round(v,4) for k,v in d["stages_ms"].items()})
PY
