# Evidence bundle of one tree, ONE gpurun call: GPU tests, smoke, bench, reference arm, launch list, ncu --set full captures.
# usage: bash tools/r02_evidence.sh <tag>     (every ncu command runs only after the same command exited 0 without ncu: the harness enforces it)
TAG=${1:-r02_v10}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
python bench.py --config5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
B="python bench.py --steps 2 --warmup 6 --latency-solves 20 --cpu-passes 1 --cpu-sample 64"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_warp -s 8 -c 1 -o gpurun_out/${TAG}_qp $B > gpurun_out/${TAG}_ncu_qp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_prepare|k_linearise|k_step_out' -s 9 -c 3 -o gpurun_out/${TAG}_prep_lin $B > gpurun_out/${TAG}_ncu_pl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_erk4 -s 1 -c 1 -o gpurun_out/${TAG}_erk4 python tools/gpu_config2.py 3 > gpurun_out/${TAG}_ncu_erk4.log 2>&1
ls -la gpurun_out | grep ${TAG}
