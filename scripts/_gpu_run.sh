cp _ab/dbl.so melissa_b200/lib/libmelissa_b200.so
timeout 900 python -m pytest tests/test_networks_gpu.py -x -q -k "record_based or table or without_dueling or flip" > gpurun_out/s37_pytest.log 2>&1; echo rc=$? >> gpurun_out/s37_pytest.log
tail -3 gpurun_out/s37_pytest.log
bash scripts/ab_bench.sh base st2 dbl
