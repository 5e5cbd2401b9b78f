#!/bin/bash
# one line per variants/libqec_*.so (and the in-tree build): scripts/time_pool2_quick.py; WIDE="0 1" also times the 512-thread build
cd "$(dirname "$0")/.."
for w in ${WIDE:-0}; do
  QEC_FUSED_WIDE=$w python scripts/time_pool2_quick.py 2>&1 | grep -v Warning
  for lib in variants/libqec_*.so; do
    QEC_FUSED_WIDE=$w QEC_LIB=$lib python scripts/time_pool2_quick.py 2>&1 | grep -v Warning
  done
done
