mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rf -s -p no:cacheprovider > gpurun_out/pytest_r2d_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2d_multi.log
tail -15 gpurun_out/pytest_r2d_multi.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --images 64 > gpurun_out/b_r2d_n2.json 2> gpurun_out/b_r2d_n2.err
tail -c 1500 gpurun_out/b_r2d_n2.json
timeout 600 $T --master-port 29512 bench.py --gpus 2 --workload collection --sets 64 --per-set 20000 --steps 2 --warmup 1 > gpurun_out/b_r2d_coll_n2.json 2> gpurun_out/b_r2d_coll_n2.err
cat gpurun_out/b_r2d_coll_n2.json
timeout 300 python bench.py --workload collection --sets 64 --per-set 20000 --steps 2 --warmup 1 > gpurun_out/b_r2d_coll_n1.json 2> gpurun_out/b_r2d_coll_n1.err
cat gpurun_out/b_r2d_coll_n1.json
tail -n 5 gpurun_out/b_r2d_n2.err
