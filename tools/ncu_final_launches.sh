set -e
# usage (on the GPU box): tools/ncu_final_launches.sh [round tag, default r02]; then here: python tools/ncu_summarise.py profiles/<tag>_final_launches.csv
TAG=${1:-r02}
export SRSB200_SUBBATCHES_DEV=1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/pre_ncu.json 2> gpurun_out/pre_ncu.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active -k regex:"scan_kernel|job_kernel|extract_kernel|emit_kernel|regroup" --clock-control none -s 72 -c 24 --csv --log-file gpurun_out/${TAG}_final_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/${TAG}_final_launches.csv | cut -c1-200
