#!/bin/bash
# build the in-tree library, then run a command on the B200 box: scripts/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m qecnetworks_b200.build > /dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
