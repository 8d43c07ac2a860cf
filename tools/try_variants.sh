#!/bin/bash
# A/B of prebuilt libraries (variants/lib_*.so) on one box: config 2 and config 3, value leg only.
# A file variants/lib_NAME.args may hold extra bench.py arguments per run, one line per run (default: one run, no extra arguments).
O=gpurun_out; mkdir -p $O
cp sushi_b200/libsushi_b200.so /tmp/lib_keep.so
for f in variants/lib_*.so; do
  name=$(basename $f .so); name=${name#lib_}
  cp $f sushi_b200/libsushi_b200.so
  argsfile=variants/lib_$name.args
  if [ -f $argsfile ]; then mapfile -t runs < $argsfile; else runs=(""); fi
  for extra in "${runs[@]}"; do
    tag=$(echo "$name $extra" | tr -s ' -' '__' | sed 's/_$//')
    for wl in config2 config3; do
      timeout 120 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-load-leg $extra > $O/var_${tag}_$wl.json 2>/dev/null
      python - "$O/var_${tag}_$wl.json" "$name [$extra] $wl" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[2], d['value'], 'events/s; match kernel', d['roofline']['kernel_ms_per_step'].get('match_fused'), 'ms; mismatches', d['shift_check']['mismatches'], 'sm MHz', d['clocks']['sm_mhz'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
    done
  done
done
cp /tmp/lib_keep.so sushi_b200/libsushi_b200.so
