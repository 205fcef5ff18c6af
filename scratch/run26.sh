mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -x -p no:cacheprovider -k "tail_kernel or profile_marks or fused_octave_cascade or write_exactly or graph_replay or extrema_kernel_forms" > gpurun_out/pytest_tail.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tail.log
tail -5 gpurun_out/pytest_tail.log | cut -c1-400
for t in 1 0; do
  SIFT_B200_TAIL=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tail$t.json 2> gpurun_out/bench_tail$t.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_tail$t.json").read().strip().splitlines()[-1])
    print("TAIL=$t 4K value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",d.get("latency"),"stages",d.get("stages_ms"),"nl",d.get("stage_launches"))
except Exception as e: print("fail",e)
PY
  SIFT_B200_TAIL=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --width 1920 --height 1080 > gpurun_out/bench_tail${t}_1080.json 2> gpurun_out/bench_tail${t}_1080.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_tail${t}_1080.json").read().strip().splitlines()[-1])
    print("TAIL=$t 1080p value",round(d["value"],1),"lat",d.get("latency"),"stages",d.get("stages_ms"))
except Exception as e: print("fail",e)
PY
done
