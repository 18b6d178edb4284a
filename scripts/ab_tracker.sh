#!/bin/bash
# A/B of prebuilt library variants (build/lib_*.so) on the tracker bench: bash scripts/ab_tracker.sh
set -e
cp ptam_cg_b200/csrc/libptam_b200.so /tmp/lib_cur.so
for v in cur "$@"; do
  if [ "$v" != cur ]; then cp build/lib_$v.so ptam_cg_b200/csrc/libptam_b200.so; fi
  python bench.py --no-ba --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['value']), round(d['e2e']['value']), {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})"
done
cp /tmp/lib_cur.so ptam_cg_b200/csrc/libptam_b200.so
