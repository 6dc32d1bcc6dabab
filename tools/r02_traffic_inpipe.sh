# DRAM bytes of one steady-state QP launch with the L2 as k_prepare / k_linearise leave it (ncu --cache-control none)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 6 --latency-solves 20 --cpu-passes 1 --cpu-sample 64"
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:'k_qp_warp|k_linearise|k_prepare' -s 24 -c 3 --csv --log-file gpurun_out/r02_v11_traffic_inpipe.csv $B > gpurun_out/r02_v11_traffic_inpipe.log 2>&1
cat gpurun_out/r02_v11_traffic_inpipe.csv | tail -14
