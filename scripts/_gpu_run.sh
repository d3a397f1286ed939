cp _ab/cf.so melissa_b200/lib/libmelissa_b200.so
timeout 900 python -m pytest tests/test_networks_gpu.py -q -k "controlling_rows or 200_nodes or large" > gpurun_out/s39_pytest.log 2>&1; echo rc=$? >> gpurun_out/s39_pytest.log
tail -8 gpurun_out/s39_pytest.log
