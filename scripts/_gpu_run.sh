python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo bench rc=$?
CMD="python bench.py --no-graph --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 2 --warmup 3"
$CMD > gpurun_out/r02f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 120 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu1.log 2>&1; echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:"conv2_attn_kernel|attn_table_mma4_kernel|gemm_bf16_tcgen05|env_round_kernel|ctrl_need_list" -s 208 -c 8 -o gpurun_out/r02f_prof -f $CMD > gpurun_out/r02f_ncu2.log 2>&1; echo ncu2 rc=$?
