mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_r2w.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2w.log
tail -6 gpurun_out/pytest_r2w.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
