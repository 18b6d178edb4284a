#!/bin/bash
# A/B of prebuilt library variants (build/lib_*.so) on the C4 bundle adjustment: bash scripts/ab_ba.sh v1 v2
cp ptam_cg_b200/csrc/libptam_b200.so /tmp/lib_cur.so
for v in cur "$@"; do
  if [ "$v" != cur ]; then cp build/lib_$v.so ptam_cg_b200/csrc/libptam_b200.so; fi
  echo "== $v"; python scripts/ba_quick.py C4
done
cp /tmp/lib_cur.so ptam_cg_b200/csrc/libptam_b200.so
