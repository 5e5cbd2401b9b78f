#!/bin/bash
# round-2 measurement batch (one gpurun call on one GPU): parity margins, benches of all configs and of the reference
# arm, launch list of the step, ncu --set full of the pool2 kernels, bit-reproducibility stress.
# Outputs under gpurun_out/r02m_*; scripts/r02_collect.sh copies the summaries to profiles/ afterwards.
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -v -s --timeout 900 > gpurun_out/r02m_tests_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_tests_v.log
python bench.py --steps 200 --warmup 5 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err
python bench.py --config 1 --steps 500 > gpurun_out/r02m_bench_c1.json 2> gpurun_out/r02m_bench_c1.err
python bench.py --config 4 --steps 50 > gpurun_out/r02m_bench_c4.json 2> gpurun_out/r02m_bench_c4.err
python bench.py --config 5 --steps 3 > gpurun_out/r02m_bench_c5.json 2> gpurun_out/r02m_bench_c5.err
QEC_GEN2_FWD_TC=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02m_bench_n1_fwdtc.json 2> /dev/null
python scripts/stress_fused.py > gpurun_out/r02m_stress.log 2>&1
python scripts/test_gen2_bwd.py 0 > gpurun_out/r02m_gen2_bwd.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02m_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu_launch.log 2>&1
python scripts/ncu_pool2.py > gpurun_out/r02m_plain_pool2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fused2|gen2_bwd_kernel" -s 8 -c 4 -o gpurun_out/prof_r02_pool2 -f \
    python scripts/ncu_pool2.py > gpurun_out/r02m_ncu_pool2.log 2>&1
tail -3 gpurun_out/r02m_tests_v.log
