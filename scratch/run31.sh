mkdir -p gpurun_out
for cfg in "1 1" "1 2" "4 1" "4 2" "8 2"; do
  set -- $cfg
  for wh in "3840 2160" "1920 1080"; do
    set -- $cfg $wh
    SIFT_B200_TAIL_TILES=$1 SIFT_B200_TAIL_CTAS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --width $3 --height $4 > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_t.json").read().strip().splitlines()[-1])
    print("TILES=$1 CTAS=$2 $3x$4 value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),"pyr",round(d["stages_ms"]["pyramid"],4),"ext",round(d["stages_ms"]["extrema"],4),d["roofline"].get("tail",{}).get("ms"))
except Exception as e: print("fail",e,open("gpurun_out/bench_t.err").read()[-300:])
PY
  done
done
