B="python bench.py --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 10 --warmup 3"
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_networks_gpu.py tests/test_policy_gpu.py -x -q > gpurun_out/s14_pytest.log 2>&1; echo rc=$? >> gpurun_out/s14_pytest.log
tail -4 gpurun_out/s14_pytest.log
timeout 300 $B > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s14_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_conv1"]["kernel_ms"], d["roofline_tensor"]["kernel_ms"], d["roofline_tensor"]["frac"])
PY
