L=gpurun_out/split_ab2.log; rm -f $L
for v in "X=1" "MGB_ATTN_LONG_MIN=0" "MGB_ATTN_SPLIT=1 MGB_ATTN_OCC3=1"; do echo "== b64 $v" >> $L; env $v timeout 120 python tools/b64_step.py 215 1 >> $L 2>&1; done
for v in "X=1" "MGB_ATTN_PF_MB=100000" "MGB_ATTN_PF_MB=32"; do echo "== long32 $v" >> $L; env $v timeout 120 python tools/long_step.py 2>&1 | grep -v "KV scan" >> $L; done
for v in "MGB_NO_ATTN_SPLIT=1" "MGB_ATTN_SPLIT=1" "MGB_ATTN_SPLIT=2" "MGB_ATTN_SPLIT=3"; do echo "== long16 $v" >> $L; env $v timeout 120 python tools/long_step.py 2400 2 16 2>&1 | grep -v "KV scan" >> $L; done
cat $L
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_final.log; tail -2 gpurun_out/bench_final.log | cut -c1-600
