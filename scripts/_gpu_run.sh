B="python bench.py --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 20 --warmup 5"
cp _ab/biasmma.so melissa_b200/lib/libmelissa_b200.so
timeout 900 python -m pytest tests/test_networks_gpu.py -x -q > gpurun_out/s30_pytest.log 2>&1; echo rc=$? >> gpurun_out/s30_pytest.log
tail -3 gpurun_out/s30_pytest.log
for i in 1 2; do for v in prep64 biasmma; do
cp _ab/$v.so melissa_b200/lib/libmelissa_b200.so
timeout 300 $B > gpurun_out/s30_bench_${v}_$i.json 2> gpurun_out/s30_bench_${v}_$i.err; echo rc=$?
done; done
python - <<'PY'
import json
for i in (1,2):
  for v in ("prep64","biasmma"):
    d=json.loads(open(f"gpurun_out/s30_bench_{v}_{i}.json").read().strip().splitlines()[-1])
    print(v, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_conv1"]["kernel_ms"], d["roofline_tensor"]["kernel_ms"])
PY
