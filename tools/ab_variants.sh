#!/bin/bash
# A/B of library variants built by tools/build_variant.sh on one box: tools/ab_variants.sh name1 name2 ...
for b in 256 ${AB_LATE:-2048}; do
for v in "$@"; do
  echo "== $v burnin $b: $(MARLLB_B200_LIB=marllb_b200/_variants/$v.so python tools/quick_ms.py --steps 100 --burnin $b 2>&1 | tail -1)"
done; done
