M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed
for l in 6 3; do
RSP_GRAPH=0 RSP_LANES=$l ncu --replay-mode app-range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2j_apprange_${l}lanes_cfg2.csv python tools/profile_chain.py --range --cpis 48 > gpurun_out/r2j_apprange_$l.log 2>&1
grep -v "^==" gpurun_out/r2j_apprange_${l}lanes_cfg2.csv | awk -F'","' 'NR>1{print $11, $13}' | tr -d '"'
done
