#!/bin/bash
# after scripts/r02_measure.sh came back: summaries -> profiles/ (tracked); the .ncu-rep files stay in gpurun_out/
cd "$(dirname "$0")/.."
set -e
G=gpurun_out; P=profiles
cp $G/r02m_tests_v.log $P/r02_gpu_tests_v.log
python scripts/margins_table.py $P/r02_gpu_tests_v.log > $P/r02_parity_margins.md
for c in n1 ref c1 c4 c5 n1_fwdtc; do cp $G/r02m_bench_$c.json $P/r02_bench_$c.json; done
cp $G/r02m_launches.csv $P/r02_launches_bench_graph.csv
cp $G/r02m_stress.log $P/r02_stress_fused.log
cp $G/r02m_gen2_bwd.log $P/r02_gen2_bwd_accuracy_timing.log
if [ -f $G/r02n_2gpu_tests.log ]; then cp $G/r02n_2gpu_tests.log $P/r02_2gpu_tests.log; fi
for n in 2 4 8; do if [ -f $G/r02n_bench_n$n.json ]; then cp $G/r02n_bench_n$n.json $P/r02_bench_n$n.json; fi; done
for f in r02_variants1.log r02_variants2.log r02_timeline1.log r02_timeline_pp1.log; do if [ -f $G/$f ]; then cp $G/$f $P/$f; fi; done
python scripts/ncu_traffic.py "qec_routing_fused_bwd=routing_fused2_bwd_kernel:$G/prof_r02_pool2.ncu-rep" \
    "qec_routing_fused_fwd=routing_fused2_fwd_kernel:$G/prof_r02_pool2.ncu-rep" "qec_gen2_bwd+=gen2_bwd_kernel:$G/prof_r02_pool2.ncu-rep" > /dev/null
python scripts/ncu_summary.py $G/prof_r02_pool2.ncu-rep > $P/r02_pool2_ncu.txt
ls -la $P | tail -30
