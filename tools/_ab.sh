python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_final.log; tail -2 gpurun_out/bench_final.log | cut -c1-300
python tools/b64_step.py 215 1 > gpurun_out/b64_after.log 2>&1; cat gpurun_out/b64_after.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/long_launches_after.csv python tools/long_step.py > gpurun_out/long_ncu.log 2>&1; tail -2 gpurun_out/long_ncu.log
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attention_kernel -c 2 -f -o gpurun_out/long_attn python tools/long_step.py > gpurun_out/long_ncu_full.log 2>&1; tail -2 gpurun_out/long_ncu_full.log; ls -la gpurun_out/*.ncu-rep
