timeout 900 python -m pytest tests/test_networks_gpu.py -x -q -k "record_based or table or without_dueling" > gpurun_out/s34_pytest.log 2>&1; echo rc=$? >> gpurun_out/s34_pytest.log
tail -12 gpurun_out/s34_pytest.log
