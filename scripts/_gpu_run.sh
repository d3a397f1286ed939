timeout 900 python -m pytest tests/test_training_gpu.py -x -q > gpurun_out/s16_pytest.log 2>&1; echo rc=$? >> gpurun_out/s16_pytest.log
tail -25 gpurun_out/s16_pytest.log
python scripts/profile_update.py 4096 > gpurun_out/s16_update.log 2>&1; grep -E "^\{|Self CUDA time" gpurun_out/s16_update.log
