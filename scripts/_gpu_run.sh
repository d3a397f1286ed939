B="python bench.py --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 20 --warmup 5"
cp _ab/c2perm.so melissa_b200/lib/libmelissa_b200.so
timeout 900 python -m pytest tests/test_networks_gpu.py -x -q -k "conv2 or bf16_forward or table" > gpurun_out/s29_pytest.log 2>&1; echo rc=$? >> gpurun_out/s29_pytest.log
tail -3 gpurun_out/s29_pytest.log
for i in 1 2; do for v in old prep64 prep40 c2perm; do
cp _ab/$v.so melissa_b200/lib/libmelissa_b200.so
timeout 300 $B > gpurun_out/s29_bench_${v}_$i.json 2> gpurun_out/s29_bench_${v}_$i.err; echo rc=$?
done; done
python - <<'PY'
import json
for i in (1,2):
  for v in ("old","prep64","prep40","c2perm"):
    d=json.loads(open(f"gpurun_out/s29_bench_{v}_{i}.json").read().strip().splitlines()[-1])
    print(v, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_conv1"]["kernel_ms"], d["roofline_tensor"]["kernel_ms"])
PY
