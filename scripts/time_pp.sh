#!/bin/bash
cd "$(dirname "$0")/.."
for lib in variants/libqec_*.so; do echo "== $lib"; QEC_LIB=$lib timeout 300 python scripts/xcheck_pp.py time 2>&1 | grep -v Warn | tail -3; done
