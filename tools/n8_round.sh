TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 280 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --batch 512 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_n8_b512.json 2> gpurun_out/r02_bench_n8_b512.err
timeout 240 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err
tail -c 1500 gpurun_out/r02_bench_n8_b512.json; echo; tail -c 600 gpurun_out/r02_bench_n8_final.json; tail -3 gpurun_out/r02_bench_n8_b512.err
