B="python bench.py --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 10 --warmup 3"
timeout 900 python -m pytest tests/test_networks_gpu.py tests/test_policy_gpu.py -x -q > gpurun_out/s22_pytest.log 2>&1; echo rc=$? >> gpurun_out/s22_pytest.log
tail -6 gpurun_out/s22_pytest.log
for hp in 1 2; do
MLS_ATTN_HP=$hp timeout 300 $B > gpurun_out/s22_bench_hp$hp.json 2> gpurun_out/s22_bench_hp$hp.err; echo rc=$?
done
python - <<'PY'
import json
for n in ("hp1","hp2"):
    try:
        d=json.loads(open(f"gpurun_out/s22_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_conv1"]["kernel_ms"], d["roofline_tensor"]["kernel_ms"])
    except Exception as e: print(n, "fail", e)
PY
