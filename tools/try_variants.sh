#!/bin/bash
# Bench the compile-time variants of the match kernel built into variants/lib_*.so (config 2 and config 3, value leg only).
O=gpurun_out; mkdir -p $O
cp sushi_b200/libsushi_b200.so /tmp/lib_keep.so
for f in variants/lib_*.so; do
  name=$(basename $f .so); name=${name#lib_}
  cp $f sushi_b200/libsushi_b200.so
  for wl in config2 config3; do
    timeout 120 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-load-leg > $O/var_${name}_$wl.json 2>/dev/null
    python - "$O/var_${name}_$wl.json" "$name $wl" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[2], d['value'], 'events/s; match kernel', d['roofline']['kernel_ms_per_step'].get('match_fused'), 'ms; mismatches', d['shift_check']['mismatches'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
  done
done
cp /tmp/lib_keep.so sushi_b200/libsushi_b200.so
