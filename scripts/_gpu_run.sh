timeout 900 python -m pytest tests/test_training_gpu.py -x -q > gpurun_out/s19_pytest.log 2>&1; echo rc=$? >> gpurun_out/s19_pytest.log
tail -6 gpurun_out/s19_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline --no-extra --no-flip > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; echo rc=$?
tail -c 1500 gpurun_out/r02f_bench_2gpu.json
