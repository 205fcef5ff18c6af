mkdir -p gpurun_out
for c in 2 3 4 5 6 8; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --contexts $c > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c.json").read().strip().splitlines()[-1])
    print("contexts=$c value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1))
except Exception as e: print("fail",e,open("gpurun_out/bench_c.err").read()[-300:])
PY
done
