mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rf --maxfail=25 -k "not 8k_set" -p no:cacheprovider > gpurun_out/pytest_r2b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2b.log
tail -4 gpurun_out/pytest_r2b.log
B="python bench.py --images 16 --steps 2 --warmup 3 --no-cpu-baseline"
for f in 0 2 3 4 5; do
SIFT_B200_EXTREMA=$f $B > gpurun_out/b_r2b_ex$f.json 2> gpurun_out/b_r2b_ex$f.err
done
python bench.py --images 64 --steps 3 --warmup 3 > gpurun_out/b_r2b_default.json 2> gpurun_out/b_r2b_default.err
SIFT_B200_GRAPH=0 $B > gpurun_out/b_r2b_nograph.json 2> gpurun_out/b_r2b_nograph.err
