mkdir -p gpurun_out /tmp/asshipped
R=${GRAFT_REPO_ROOT:-/root/repo}
( cd /tmp/asshipped; TIMEFORMAT=%R; time $R/oracle/_ref/sift $R/tests/golden/image1.png $R/tests/golden/image2.png > out.txt 2>&1 ) 2> /tmp/asshipped/wall.txt &
timeout 900 python -m pytest tests -m gpu -q -rf --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_r2c.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2c.log
tail -4 gpurun_out/pytest_r2c.log
B="python bench.py --images 16 --steps 2 --warmup 3 --no-cpu-baseline"
for f in 0 2 3 4; do
SIFT_B200_EXTREMA=$f $B > gpurun_out/b_r2c_ex$f.json 2> gpurun_out/b_r2c_ex$f.err
done
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_extrema4|k_stream|k_orient|k_refine" -s 40 -c 12 -o gpurun_out/prof_r2c python scratch/one_detect.py 3 > gpurun_out/ncu_r2c.log 2>&1
tail -3 gpurun_out/ncu_r2c.log
wait
python - <<'PY'
import json, os
wall = float(open("/tmp/asshipped/wall.txt").read().strip().splitlines()[-1])
out = open("/tmp/asshipped/out.txt").read()
fin = [l for l in out.splitlines() if l.startswith("Final keypoints")]
cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
rec = {"command": "oracle/_ref/sift image1.png image2.png (reference main.cpp + sift.cpp exactly as shipped, g++ -O3)",
       "wall_s": wall, "host_cores": os.cpu_count(), "cpu": cpu[0] if cpu else "?", "threads_used": 1,
       "stdout_final": fin, "concurrent_load": "the GPU tests / benchmarks of this call on other cores"}
json.dump(rec, open("gpurun_out/r2_config1_asshipped.json", "w"), indent=1)
print(rec)
PY
