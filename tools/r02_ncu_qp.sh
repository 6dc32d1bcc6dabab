# one steady-state k_qp_warp launch of the bench (config 3) under ncu --set full   usage: bash tools/r02_ncu_qp.sh <tag> [skip]
TAG=${1:-r02}; SKIP=${2:-8}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 6 --latency-solves 20 --cpu-passes 1 --cpu-sample 64"
ncu --set full --clock-control none --import-source on -k regex:k_qp_warp -s $SKIP -c 1 -o gpurun_out/${TAG}_qp $B > gpurun_out/${TAG}_ncu_qp.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_qp.log
