mkdir -p gpurun_out
for f in 300000 600000 2500000 9000000; do
  for wh in "3840 2160" "1920 1080"; do
    set -- $wh
    SIFT_B200_FORK_MIN_PX=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --width $1 --height $2 > gpurun_out/bench_fork.json 2> gpurun_out/bench_fork.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_fork.json").read().strip().splitlines()[-1])
    print("FORK=$f $1x$2 value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4))
except Exception as e: print("fail",e)
PY
  done
done
