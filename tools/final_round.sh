set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/final_pytest.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cp easynlp_b200/lib/libclipk.sha256 gpurun_out/r02_launches.csv.sha256
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 200 python tests/preprocess_bench.py 256 640 480 > gpurun_out/preprocess_bench.log 2>&1
tail -3 gpurun_out/final_pytest.log; cat gpurun_out/bench_final.json | head -c 3000; tail -2 gpurun_out/bench_ref.json | head -c 800; cat gpurun_out/preprocess_bench.log
