python bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_b1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu_b1.log; tail -2 gpurun_out/bench_2gpu_b1.log | cut -c1-200
python bench.py --gpus 2 --batch 64 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_b64.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu_b64.log; tail -2 gpurun_out/bench_2gpu_b64.log | cut -c1-200
