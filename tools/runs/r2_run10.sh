mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_dist_nccl.py tests/test_exp_shim.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2/pytest_r10.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline --no-ragged 2>gpurun_out/r2/b10_1.err | tail -1 > gpurun_out/r2/bench_r10_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b10_2.err | tail -1 > gpurun_out/r2/bench_r10_n2_dpgraph.json
CWN_BENCH_DP_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b10_3.err | tail -1 > gpurun_out/r2/bench_r10_n2_3piece.json
tail -6 gpurun_out/r2/pytest_r10.log; grep -i "capturing\|error" gpurun_out/r2/b10_2.err | head -5
