L=gpurun_out/split_ab3.log; rm -f $L
for v in "X=1" "MGB_ATTN_OCC3=1"; do echo "== b64 $v" >> $L; env $v timeout 120 python tools/b64_step.py 215 1 >> $L 2>&1; done
for v in "X=1" "MGB_ATTN_OCC3=1" "MGB_ATTN_OCC3=1 MGB_ATTN_SPLIT=2"; do echo "== long32 $v" >> $L; env $v timeout 120 python tools/long_step.py 2>&1 | grep -v "KV scan" >> $L; done
for v in "X=1" "MGB_ATTN_OCC3=1" "MGB_ATTN_OCC3=1 MGB_ATTN_SPLIT=3"; do echo "== long16 $v" >> $L; env $v timeout 120 python tools/long_step.py 2400 2 16 2>&1 | grep -v "KV scan" >> $L; done
cat $L
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "long_kv" 2>&1 | tail -3
