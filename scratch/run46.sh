mkdir -p gpurun_out
SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_old.log 2>&1
timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_new.log 2>&1
cmp gpurun_out/ab_old.log gpurun_out/ab_new.log && echo "AB IDENTICAL"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "stage_lists or orientation_stage or descriptor_stage or output_order or bit_reproducible or graph_replay or small_and_odd or flat_image or profile_marks or batch_detect or contexts_come" 2>&1 | tail -3
for lib in prev new prev new; do
  if [ $lib = new ]; then unset SIFT_B200_LIB; else export SIFT_B200_LIB=$PWD/scratch/variants/libsift_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
    print("$lib value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),"refine",round(d["stages_ms"]["refine"],4),"orient",round(d["stages_ms"]["orient"],4),"sort",round(d["stages_ms"]["sort"],4), d["stage_launches"]["sort"])
except Exception as e: print("$lib fail", e, open("gpurun_out/bench_ab.err").read()[-400:])
PY
done
