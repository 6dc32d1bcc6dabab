# multi-GPU bench lines (torchrun, one rank per GPU)   usage: bash tools/r02_multi.sh <tag> <N> [<N> ...]
TAG=$1; shift
mkdir -p gpurun_out
for N in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  echo "N=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
    print("N=%d value %.4g e2e %.4g ms %.3f p99 %.3f workload %s"%(d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["latency_ms"]["p99"], d["config"]["workload"][:60]))
    for k in ("config5","config5_feasible_start"):
        c=d.get(k)
        if c: print("  ",k,c["instances"],"%.3f s"%c["seconds"],"sqp it/s %.4g"%c["sqp_iterations_per_s"],"conv %.3f"%c["converged_frac"])
except Exception as e: print("parse failed", e)
PY
done
