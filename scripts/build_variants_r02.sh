#!/bin/bash
# round-2 timing variants of the packed cluster kernels (scripts/build_variant.sh NAME flags...), timed in ONE gpurun
# call by scripts/time_variants.sh
set -e
cd "$(dirname "$0")/.."
python -m qecnetworks_b200.build > /dev/null
scripts/build_variant.sh base  -DQEC_SWP=0 -DQEC_ESTRIN=0 -DQEC_SPREAD=0 &
scripts/build_variant.sh swp   -DQEC_SWP=1 -DQEC_ESTRIN=0 -DQEC_SPREAD=0 &
scripts/build_variant.sh swpe  -DQEC_SWP=1 -DQEC_ESTRIN=1 -DQEC_SPREAD=0 &
scripts/build_variant.sh swpes -DQEC_SWP=1 -DQEC_ESTRIN=1 -DQEC_SPREAD=1 &
wait
ls -la variants/*.so
