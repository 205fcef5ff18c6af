mkdir -p gpurun_out
# full GPU suite on the final build
timeout 1200 python -m pytest tests -m gpu -q -rf --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_r2g.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2g.log
tail -4 gpurun_out/pytest_r2g.log
# the driver's default line and the other named resolutions
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/b_r2g_default.json 2> gpurun_out/b_r2g_default.err
timeout 300 python bench.py --width 1920 --height 1080 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2g_1080p.json 2> gpurun_out/b_r2g_1080p.err
timeout 300 python bench.py --width 7680 --height 4320 --images 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_r2g_8k.json 2> gpurun_out/b_r2g_8k.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_r2g_reference.json 2> gpurun_out/b_r2g_reference.err
# launch list of one image + DRAM traffic of the pyramid kernels + full ncu of the top kernels
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 84 -c 42 --csv --log-file gpurun_out/launches_r2g.csv python scratch/one_detect.py 3 > gpurun_out/ncu_r2g_launches.log 2>&1
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_extrema4|k_stream|k_input_u8" -s 18 -c 5 -o gpurun_out/prof_r2g python scratch/one_detect.py 3 > gpurun_out/ncu_r2g.log 2>&1
tail -2 gpurun_out/ncu_r2g.log
