#!/bin/bash
for rep in 1 2; do
for v in base perm both; do
  echo "== $v"; MARLLB_B200_LIB=marllb_b200/_variants/$v.so python tools/quick_ms.py --steps 100 2>&1 | tail -1
done; done
