#!/bin/bash
# the mel A/B: sparse banded FMA vs tcgen05.mma tf32 x3 (tools/ubench/mel_umma.cu); plain run, then ncu counters of both kernels
mkdir -p gpurun_out
cd tools/ubench
for m in 80 128; do timeout 60 ./mel_umma $m 8 | tee ../../gpurun_out/mel_ab_$m.txt; echo "exit $?"; done
if [ -n "$WITH_NCU" ]; then
timeout 60 ./mel_umma 80 8 > ../../gpurun_out/mel_ab_plain.log 2>&1 &&
timeout 300 ncu --clock-control none -s 3 -c 1 -k regex:mel_sparse --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --csv --log-file ../../gpurun_out/mel_ab_ncu_sparse.csv ./mel_umma 80 8 > /dev/null 2>&1
timeout 300 ncu --clock-control none -s 3 -c 1 -k regex:mel_umma --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --csv --log-file ../../gpurun_out/mel_ab_ncu_umma.csv ./mel_umma 80 8 > /dev/null 2>&1
echo "ncu exit $?"
fi
