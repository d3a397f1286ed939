cp _ab/prepb_gx.so melissa_b200/lib/libmelissa_b200.so
timeout 900 python -m pytest tests/test_networks_gpu.py -x -q > gpurun_out/s32_pytest.log 2>&1; echo rc=$? >> gpurun_out/s32_pytest.log
tail -3 gpurun_out/s32_pytest.log
bash scripts/ab_bench.sh split prepb prepb_gx
