python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r02i_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02i_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo bench rc=$?
CMD="python bench.py --no-graph --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 2 --warmup 3"
$CMD > gpurun_out/r02i_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 570 -c 120 --csv --log-file gpurun_out/r02i_launches.csv $CMD > gpurun_out/r02i_ncu1.log 2>&1; echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:"conv2_attn_kernel|attn_table_rows_kernel|attn_table_prep_kernel|gemm_bf16_tcgen05|env_round_kernel|ctrl_need_list" -s 234 -c 9 -o gpurun_out/r02i_prof -f $CMD > gpurun_out/r02i_ncu2.log 2>&1; echo ncu2 rc=$?
