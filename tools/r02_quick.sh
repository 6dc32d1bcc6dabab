# quick GPU check of a kernel change: GPU tests + one bench line   usage: bash tools/r02_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
python bench.py --latency-solves 50 --cpu-passes 1 --cpu-sample 256 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_gpu_tests.log
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g e2e %.4g ms/step %.4f phases %s k_ipm %.3f c4 %.4g"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["phase_ms"], d["k_ipm_mean"], d["config4_one_gpu"]["value"]))
PY
