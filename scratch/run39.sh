mkdir -p gpurun_out
python scratch/one_detect.py 3 > gpurun_out/one_detect.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_refine" -s 1 -c 1 -o gpurun_out/prof_refine_cubes python scratch/one_detect.py 3 > gpurun_out/ncu_ref1.log 2>&1
SIFT_B200_CUBES=0 ncu --set full --clock-control none --import-source on -k regex:"k_refine" -s 1 -c 1 -o gpurun_out/prof_refine_nocubes python scratch/one_detect.py 3 > gpurun_out/ncu_ref0.log 2>&1
tail -1 gpurun_out/ncu_ref0.log
