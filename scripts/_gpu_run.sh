python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --no-cpu-baseline --no-extra --no-flip > gpurun_out/r02g_bench_8gpu.json 2> gpurun_out/r02g_bench_8gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench_8gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
t=d["train"]; print({k:t[k] for k in ["value","ms_per_step","update_ms","allreduce_us","allreduce_busbw_GBs","allreduce_share_of_step","weights_identical_across_ranks"]})
PY
