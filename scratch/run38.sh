mkdir -p gpurun_out
for i in 1 2; do
for lib in new noatom; do
  if [ $lib = noatom ]; then export SIFT_B200_LIB=$PWD/scratch/variants/libsift_noatom.so; else unset SIFT_B200_LIB; fi
  SIFT_B200_CUBES=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
    print("$lib ext",round(d["stages_ms"]["extrema"],4), [round(p["ms"],4) for p in d["roofline"]["per_kernel"]])
except Exception as e: print("$lib fail", e, open("gpurun_out/bench_ab.err").read()[-400:])
PY
done
done
