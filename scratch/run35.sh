mkdir -p gpurun_out
SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_old.log 2>&1
timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_new.log 2>&1
tail -1 gpurun_out/ab_old.log; tail -1 gpurun_out/ab_new.log
cmp gpurun_out/ab_old.log gpurun_out/ab_new.log && echo "AB IDENTICAL"
for i in 1 2; do
for lib in old new; do
  if [ $lib = old ]; then export SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so; else unset SIFT_B200_LIB; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print("$lib value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),"desc",round(d["stages_ms"]["describe"],4),"orient",round(d["stages_ms"]["orient"],4))
PY
done
done
