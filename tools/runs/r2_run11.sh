mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_dist_nccl.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2/pytest_r11.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b11_2.err | tail -1 > gpurun_out/r2/bench_r11_n2_overlap.json
CWN_B200_DP_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b11_3.err | tail -1 > gpurun_out/r2/bench_r11_n2_single.json
tail -12 gpurun_out/r2/pytest_r11.log; grep -i "capturing\|error" gpurun_out/r2/b11_2.err | head -5
