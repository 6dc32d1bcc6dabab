set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --latency-solves 20 --cpu-passes 1 --cpu-sample 64"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv $B > gpurun_out/r02e_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_warp -s 1 -c 1 -o gpurun_out/r02e_qp $B > gpurun_out/r02e_ncu_qp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_prepare|k_linearise' -s 2 -c 2 -o gpurun_out/r02e_prep_lin $B > gpurun_out/r02e_ncu_pl.log 2>&1
ls -la gpurun_out
