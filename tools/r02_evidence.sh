set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02e_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02e_smoke.log
python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --latency-solves 0 > gpurun_out/r02e_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_warp -s 1 -c 1 -o gpurun_out/r02e_qp python bench.py --steps 2 --warmup 3 --latency-solves 0 > gpurun_out/r02e_ncu_qp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_prepare|k_linearise' -s 2 -c 2 -o gpurun_out/r02e_prep_lin python bench.py --steps 2 --warmup 3 --latency-solves 0 > gpurun_out/r02e_ncu_pl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_erk4 -s 1 -c 1 -o gpurun_out/r02e_erk4 python tools/gpu_config2.py 3 > gpurun_out/r02e_ncu_erk4.log 2>&1
ls -la gpurun_out
