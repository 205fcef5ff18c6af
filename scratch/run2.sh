mkdir -p gpurun_out
# config 1 as shipped (CPU only, one thread, ~6 min): runs beside everything else in this call
( mkdir -p /tmp/asshipped && cd /tmp/asshipped && s=$(date +%s.%N) && $GRAFT_REPO_ROOT/oracle/_ref/sift $GRAFT_REPO_ROOT/tests/golden/image1.png $GRAFT_REPO_ROOT/tests/golden/image2.png > sift.out 2>&1; rc=$?; e=$(date +%s.%N); nproc=$(nproc); model=$(grep -m1 "model name" /proc/cpuinfo | cut -d: -f2 | xargs); fin=$(grep "Final keypoints" sift.out | tr '\n' ' '); echo "{\"command\": \"oracle/_ref/sift image1.png image2.png (reference main.cpp + sift.cpp exactly as shipped, g++ -O3)\", \"wall_s\": $(echo "$e - $s" | bc), \"exit\": $rc, \"host_cores\": $nproc, \"cpu\": \"$model\", \"threads_used\": 1, \"stdout_final\": \"$fin\", \"concurrent_load\": \"GPU tests and benchmarks of this repo on the other cores\"}" > $GRAFT_REPO_ROOT/gpurun_out/r2_config1_asshipped.json ) &
timeout 900 python -m pytest tests -m gpu -q -rf --maxfail=25 -k "not 8k_set" -p no:cacheprovider > gpurun_out/pytest_r2b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2b.log
tail -4 gpurun_out/pytest_r2b.log
B="python bench.py --images 16 --steps 2 --warmup 3 --no-cpu-baseline"
for f in 0 2 3 4 5; do
SIFT_B200_EXTREMA=$f $B > gpurun_out/b_r2b_ex$f.json 2> gpurun_out/b_r2b_ex$f.err
done
python bench.py --images 64 --steps 3 --warmup 3 > gpurun_out/b_r2b_default.json 2> gpurun_out/b_r2b_default.err
SIFT_B200_GRAPH=0 $B > gpurun_out/b_r2b_nograph.json 2> gpurun_out/b_r2b_nograph.err
wait
cat gpurun_out/r2_config1_asshipped.json
./scratch/tmem_bw > gpurun_out/tmem_bw_r2.txt 2>&1
cat gpurun_out/tmem_bw_r2.txt
