#!/bin/bash
# Run under gpurun (1 GPU): one `ncu --set full` capture of the dominant kernel
# of every bench.py workload, on the exact launch shapes bench.py times.
#   gpurun --timeout 1500 -- 'bash benchmarks/ncu_traffic.sh'
# then here:  python benchmarks/ncu_traffic.py   (writes profiles/r2_traffic.json)
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$NCU -k regex:det_tma_kernel -s 3 -c 1 -f -o gpurun_out/r2_ncu_rmse_acc $B --workloads none > gpurun_out/ncu_rmse_acc.log 2>&1
$NCU -k "regex:ens_pair_kernel|ens_metrics_kernel" -s 3 -c 1 -f -o gpurun_out/r2_ncu_crps_sweep $B --workloads crps > gpurun_out/ncu_crps.log 2>&1
$NCU -k regex:regrid -s 3 -c 1 -f -o gpurun_out/r2_ncu_regrid $B --workloads regrid > gpurun_out/ncu_regrid.log 2>&1
$NCU -k regex:spectrum_pfa_kernel -s 3 -c 1 -f -o gpurun_out/r2_ncu_spectrum_sweep $B --workloads spectrum > gpurun_out/ncu_spectrum.log 2>&1
$NCU -k regex:spectrum_pfa_kernel -s 7 -c 1 -f -o gpurun_out/r2_ncu_spectrum_latsum $B --workloads spectrum > gpurun_out/ncu_latsum.log 2>&1
# launch list of the whole default command (shares of the step, cold-cache / serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
