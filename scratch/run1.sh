mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --maxfail=25 -k "not 8k_set" -p no:cacheprovider > gpurun_out/pytest_r2a.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2a.log
B="python bench.py --images 64 --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/b_r2a_default.json 2> gpurun_out/b_r2a_default.err
SIFT_B200_GRAPH=0 $B > gpurun_out/b_r2a_nograph.json 2> gpurun_out/b_r2a_nograph.err
SIFT_B200_EXTREMA=1 $B > gpurun_out/b_r2a_ex1.json 2> gpurun_out/b_r2a_ex1.err
SIFT_B200_GRAPH=0 SIFT_B200_EXTREMA=1 SIFT_B200_CENTER=0 $B > gpurun_out/b_r2a_r1like.json 2> gpurun_out/b_r2a_r1like.err
tail -5 gpurun_out/pytest_r2a.log
