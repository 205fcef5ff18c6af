mkdir -p gpurun_out
SIFT_B200_LIB=$PWD/scratch/variants/libsift_old.so timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_old.log 2>&1
timeout 300 python scratch/ab_bytes.py > gpurun_out/ab_new.log 2>&1
tail -1 gpurun_out/ab_old.log; tail -1 gpurun_out/ab_new.log
cmp gpurun_out/ab_old.log gpurun_out/ab_new.log && echo "AB IDENTICAL"
timeout 1500 python -m pytest tests -m gpu -q -rf --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_r2z.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2z.log
tail -4 gpurun_out/pytest_r2z.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
