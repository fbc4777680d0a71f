#!/bin/bash
# per-launch DRAM traffic / tensor-pipe / L2 metrics of two steady-state steps of the final round-2 tree (a metrics
# pass, not --set full: the whole-step full report of call H was 113 MB, more than gpurun brings back)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --sustained-s 0"
timeout 420 ncu --metrics $M --clock-control none -s 700 -c 130 --csv --log-file gpurun_out/r02r_ncu_metrics.csv $B > /dev/null 2> gpurun_out/r02r_ncu.err
echo "ncu rc=$?"; wc -l gpurun_out/r02r_ncu_metrics.csv; tail -2 gpurun_out/r02r_ncu.err; du -sh gpurun_out | tail -1
