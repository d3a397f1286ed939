timeout 900 python -m pytest tests/test_networks_gpu.py tests/test_training_gpu.py -x -q > gpurun_out/s17_pytest.log 2>&1; echo rc=$? >> gpurun_out/s17_pytest.log
tail -15 gpurun_out/s17_pytest.log
