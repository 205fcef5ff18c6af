mkdir -p gpurun_out
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_orient|k_refine|k_extrema4|k_stream|k_input_u8" -s 15 -c 4 -o gpurun_out/prof_r2z_a python scratch/one_detect.py 3 > gpurun_out/ncu_r2z_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_orient|k_refine" -s 3 -c 3 -o gpurun_out/prof_r2z_b python scratch/one_detect.py 3 > gpurun_out/ncu_r2z_b.log 2>&1
tail -1 gpurun_out/ncu_r2z_b.log
