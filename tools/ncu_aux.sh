set -e
# launch list of the kernels around the decoder on a 256-TB submission (downlink source, then uplink source)
python tools/tb_breakdown.py 256 > gpurun_out/pre_ncu_aux.json 2> gpurun_out/pre_ncu_aux.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active -k regex:"rm_rx_kernel|ulsch_deint_kernel|crc_bytes_kernel|gather_copy_kernel|extract_kernel|emit_kernel" --clock-control none -c 40 --csv --log-file gpurun_out/r01_aux_launches.csv python tools/tb_breakdown.py 256 > gpurun_out/ncu_aux_run.log 2>&1
tail -2 gpurun_out/r01_aux_launches.csv | cut -c1-160
