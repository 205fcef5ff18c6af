mkdir -p gpurun_out
timeout 120 python scratch/small_detect.py > gpurun_out/small_plain.log 2>&1 && timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/small_detect.py > gpurun_out/memcheck_r2.log 2>&1
echo "exit $?" >> gpurun_out/memcheck_r2.log
tail -3 gpurun_out/small_plain.log; tail -15 gpurun_out/memcheck_r2.log
