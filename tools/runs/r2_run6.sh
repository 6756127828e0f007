mkdir -p gpurun_out/r2
timeout 300 python bench.py --steps 200 --warmup 10 --no-sweep --no-cpu-baseline 2>gpurun_out/r2/b_tc5.err | tail -1 > gpurun_out/r2/bench_tc5.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc5 -s 60 -c 8 -o gpurun_out/r2/ncu_tc5 python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline > gpurun_out/r2/ncu_tc5.log 2>&1
tail -3 gpurun_out/r2/ncu_tc5.log
