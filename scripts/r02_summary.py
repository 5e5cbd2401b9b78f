"""profiles/r02_summary.md from the files scripts/r02_measure.sh brought back (run after scripts/r02_collect.sh)"""
import json, subprocess, os, re
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def J(c):
    return json.loads(open('gpurun_out/r02m_bench_%s.json' % c).read().strip().splitlines()[-1])
n1, ref, c1, c4, c5, ftc = (J(c) for c in ("n1", "ref", "c1", "c4", "c5", "n1_fwdtc"))
step = subprocess.run(["python", "scripts/launch_step.py", "gpurun_out/r02m_launches.csv"], capture_output=True, text=True).stdout
nlaunch = int(re.search(r"launches in one step: (\d+)", step).group(1))
ncu = subprocess.run(["python", "scripts/ncu_summary.py", "gpurun_out/prof_r02_pool2.ncu-rep"], capture_output=True, text=True).stdout
def bound_of(k):
    ht, ft = k["alg_bytes"] / 6535.4e9, k["alg_flop"] / (n1["fp32_peak_tflops"] * 1e12)
    return ("hbm", k["hbm_frac"]) if ht >= ft else ("fp32", k["fp32_frac"])
ks = "\n".join("| `%s` | %.4f | %.1f %% | %s | %.3f | %s |" % ((k["kernel"], k["ms_per_step"], 100 * k["share_of_step"]) + bound_of(k) + ("x".join(str(x) for x in k["dims"]),))
               for k in n1["kernels"][:12])
r = n1["roofline"]; r2 = n1.get("roofline_second", {})
hbm_frac = r.get("frac_hbm", r.get("other_roof_frac", float("nan")))
def N(n):
    return json.loads(open('gpurun_out/r02n_bench_n%d.json' % n).read().strip().splitlines()[-1])
n2, n4, n8 = N(2), N(4), N(8)
out = f"""# Round 2 profiles (B200, sm_100a) -- index

All captures: `gpurun` on a fresh B200, the profiled command first run to exit 0 without ncu.  Bench numbers are never
taken under a profiler.  One measurement batch (`scripts/r02_measure.sh`) produced the final files; `scripts/r02_collect.sh`
copied the summaries here.

| file | what |
|---|---|
| `r02_bench_n1.json` | **final** `python bench.py --steps 200 --warmup 5` (N = 1, BASELINE `configs[1]`, CUDA graph, incl. `cpu_baseline`, `roofline`, `extra.roofline_second`, pipelined `e2e`): **{n1['value']:.0f} samples/s, {n1['ms_per_step']:.3f} ms/step** (e2e {n1['e2e']['value']:.0f}) |
| `r02_bench_ref.json` | `python bench.py --impl reference` (oracle port of the reference CPU path, 16 host cores, B = 8 per step): {ref['value']:.2f} samples/s |
| `r02_bench_c1.json` | `--config 1` (one `QecModule` layer forward, B = 8): {c1['value']:.0f} samples/s; `latency_us`: layer under a CUDA graph {c1['latency_us']['layer_graph_replay']:.1f} us (roofline {c1['latency_us']['roofline_us']:.1f} us), with device input copy {c1['latency_us']['with_device_input_copy']:.1f}, from host {c1['latency_us']['end_to_end_from_host']:.1f} |
| `r02_bench_c4.json` | `--config 4` (Siamese `QecSiaNet` training step, B = 16 pairs, spread x2 + 0.1 pose-difference loss in one launch): {c4['value']:.0f} pairs/s, {c4['ms_per_step']:.3f} ms/step; relative-rotation error of the predicted class poses in `extra` |
| `r02_bench_c5.json` | `--config 5` (stress sweep, inference, B = 1024; a step = one corner of K {{9,64}} x O {{32,512}} x T {{1,10}}): mean {c5['ms_per_step']:.2f} ms per corner, per-corner times in `extra` |
| `r02_bench_n1_fwdtc.json` | the headline config with the package's own tcgen05 forward of `quater_gen2` (`QEC_GEN2_FWD_TC=1`) instead of cuBLAS: {ftc['ms_per_step']:.3f} ms/step -- slower, hence opt-in |
| `r02_bench_n2.json`, `r02_bench_n4.json`, `r02_bench_n8.json` | final code under torchrun, weak scaling (B = 32 per GPU, all-reduce inside the graph): N = 2 {n2['value']:.0f} samples/s ({n2['ms_per_step']:.3f} ms/step), N = 4 {n4['value']:.0f} ({n4['ms_per_step']:.3f}), N = 8 **{n8['value']:.0f} samples/s** ({n8['ms_per_step']:.3f} ms/step) = {n8['value']/n1['value']:.2f} x of N = 1 (efficiency {n2['value']/n1['value']/2:.3f} / {n4['value']/n1['value']/4:.3f} / {n8['value']/n1['value']/8:.3f}) |
| `r02_launches_bench_graph.csv` | launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (`ncu --metrics gpu__time_duration.sum --clock-control none -c 900`), final code |
| `r02_pool2_ncu.txt` | `ncu --set full --clock-control none --import-source on -k regex:"fused2\\|gen2_bwd_kernel" -s 8 -c 4 ... python scripts/ncu_pool2.py`: key metrics + top stalls of the three pool2 kernels of the training step (routing forward, routing backward, tcgen05 backward of `quater_gen2`, both products) |
| `ncu_traffic.json` | DRAM read/write bytes per launch from that capture (feeds `roofline.traffic` in `bench.py`) |
| `r02_sass_tcgen05.txt` | `cuobjdump -sass` evidence of the tensor-core kernel: `UTCHMMA` (tcgen05.mma), `UTMALDG` (TMA), `LDTM` (tcgen05.ld), `UTCBAR`, `SYNCS` |
| `r02_gpu_tests_v.log`, `r02_parity_margins.md` | `pytest -m gpu -v -s` of the final code and the table of every `err / tol` it prints |
| `r02_2gpu_tests.log` | the tests that need two GPUs (`gpurun --gpus 2`): NCCL data-parallel step, `nn.DataParallel` against one device, module on `cuda:1` |
| `r02_race_hunt.log` | the race in the routing backward: before / after (`scripts/race_hunt_bwd.py`, `scripts/flaky_hunt.py`) |
| `r02_stress_fused.log` | `scripts/stress_fused.py` (bit-reproducibility, 30 repetitions x 7 geometries) |
| `r02_gen2_bwd_accuracy_timing.log` | `scripts/test_gen2_bwd.py`: `qec_gen2_bwd` against float64 and its timing with MMAs / split compiled out |
| `r02_variants1.log`, `r02_variants2.log` | A/B timing of compile-time variants of the cluster kernels (`QEC_SWP` bits, Estrin, spread, shuffle reduce; wide build) |
| `r02_timeline1.log`, `r02_timeline_pp1.log` | `clock64` phase timelines of one capsule: second generation (base / software-pipelined), third generation (`routing_fused3.cu`) |

## Headline (BASELINE `configs[1]`, N = 1)

{n1['ms_per_step']:.3f} ms/step, {n1['value']:.0f} samples/s device-resident, {n1['e2e']['value']:.0f} end to end (H2D {n1['e2e']['h2d_bytes_per_step']} B + loss D2H every step);
{n1['gpu_launches_per_step']} launches of this package per step, {nlaunch} launches in total (launch list).  Clocks {n1['clocks']['sm_mhz']:.0f} / {n1['clocks']['sm_max_mhz']:.0f} MHz, reasons {n1['clocks']['reasons']}.
CPU arm {n1['cpu_baseline']['value']:.2f} samples/s ({n1['cpu_baseline']['cores']} cores, `{n1['cpu_baseline']['kind']}`).

`roofline` (dominant kernel `{r['kernel']}`, {r['ms']:.3f} ms): bound `{r['bound']}`, achieved {r['achieved']:.1f} {r['unit']} of {r['peak']:.1f} = **{r['frac']:.3f}**;
on the HBM roof with the SURVEY 8(d) bytes ({r['alg_bytes']/1e9:.3f} GB): {hbm_frac:.3f}; measured DRAM traffic {r['traffic']/1e9:.3f} GB = {r['traffic_over_alg']:.2f} x the algorithmic bytes.
Second kernel `{r2.get('kernel')}`: {r2.get('ms', 0):.3f} ms, `{r2.get('bound')}` {r2.get('frac', 0):.3f}.

Kernel times inside the step (`bench.py`'s own CUDA-event pass, warm, eager; `kernels[]` of `r02_bench_n1.json`):

| entry point | ms per step | share | bound | fraction of that roof | dims |
|---|---|---|---|---|---|
{ks}

## One training step, by kernel (launch list)

{step}
The shares agree with the event pass above (backward 45 / 46 %, forward 20 / 21 %, tcgen05 backward 13 / 14 %).
ATen launches left ({nlaunch - n1['gpu_launches_per_step'] - 1}): fills, 1 add (the two gradient contributions to the pool2 input poses), relu + its backward,
the loss scale, one copy; library GEMM left: the `quater_gen2` forward (`cutlass3x_sm100_tensorop_s256x256x8tf32gemm`, 120 us, 86 % of the
HBM write roof).  Round 1 ended with 62 launches, 18 own.

## `ncu --set full` of the pool2 kernels (1 280 capsules x 32 768 votes, B = 32)

```
{ncu}```

Reading: the routing kernels are far from HBM (DRAM 17-19 % busy, `traffic / algorithmic = 1.04` for the backward: nothing re-read,
nothing spilled) and not saturated on the fma pipe over the whole kernel; the shared-memory pipe
(`l1tex__data_pipe_lsu_wavefronts`) is the busiest unit (54 % / 48 % of elapsed, saturated while both CTAs of an SM sweep) and
`warps_active` 22 % shows the serial chains (7 of 8 warps parked while one solves).  `gen2_bwd_kernel`: tensor pipe 14 % busy,
L2 -> SM 6 TB/s, DRAM 3.7 TB/s: streaming, see DESIGN.md section 3 for what bounds it.

## What was tried on the cluster kernels this round (all measured, `r02_variants*.log`, `r02_timeline*.log`)

| variant | forward ms | backward ms (plain / pre-split) | verdict |
|---|---|---|---|
| base (round-1 final) | 0.505 | 1.136 / 1.188 | - |
| chains compiled out (`QEC_PROBE_NOCHAIN`, wrong results, timing probe) | 0.240 | - / 0.717 | the chains cost 52 % / 40 % |
| software-pipelined shared-memory loads, forward routing sweeps (`QEC_SWP=1`) | 0.508 | = | neutral |
| ... forward sweep 0 (`=2`), epilogue (`=4`), all forward (`=7`) | 0.520 / 0.506 / 0.514 | = | neutral / slower |
| ... backward accumulation sweeps (`=8`), backward sweep 0 (`=16`) | = | 1.180 / 1.236, 1.135 / 1.188 | slower / neutral |
| Estrin distance polynomial + SWP | 0.528 | 1.206 / 1.262 | slower |
| capsule order spread over clusters + Estrin + SWP | 0.532 | 1.261 / 1.324 | slower |
| 512-thread build (`wide=1`) | 0.518 | 1.127 / 1.182 | forward slower, backward -0.8 % (kept opt-in) |
| third generation, two capsules per cluster, solver warps, named barriers (`routing_fused3.cu`) | 0.553 | - | slower; 15 instead of 16.5 clusters co-resident |
"""
open('profiles/r02_summary.md', 'w').write(out)
print(out[:3000])
