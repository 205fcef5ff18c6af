mkdir -p gpurun_out
SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_old.log 2>&1
timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_new.log 2>&1
tail -1 gpurun_out/ab_old.log; tail -1 gpurun_out/ab_new.log
cmp gpurun_out/ab_old.log gpurun_out/ab_new.log && echo "AB IDENTICAL"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "stage_lists or extrema_kernel_forms or write_exactly or tail_kernel or graph_replay or window_and_bin or small_and_odd or other_interval" 2>&1 | tail -3
for i in 1 2; do
for cub in 0 1; do
  SIFT_B200_CUBES=$cub timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print("cubes=$cub value",round(d["value"],1),"lat",round(d["latency"]["ms_per_image_one_stream"],4),"ext",round(d["stages_ms"]["extrema"],4),"refine",round(d["stages_ms"]["refine"],4))
PY
done
done
