mkdir -p gpurun_out
for lib in new ref8 ref10; do
for cub in 0 1; do
  if [ $lib = new ]; then unset SIFT_B200_LIB; else export SIFT_B200_LIB=$PWD/scratch/variants/libsift_$lib.so; fi
  SIFT_B200_CUBES=$cub timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
    print("$lib cubes=$cub value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),"ext",round(d["stages_ms"]["extrema"],4),"refine",round(d["stages_ms"]["refine"],4))
except Exception as e: print("$lib fail", e, open("gpurun_out/bench_ab.err").read()[-400:])
PY
done
done
