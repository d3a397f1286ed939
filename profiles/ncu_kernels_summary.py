"""Key counters of every kernel in one `ncu --set full` report -> profiles/r02_ncu_kernels.txt + profiles/r02_traffic.json.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python profiles/ncu_kernels_summary.py raw.csv"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1]))); h = rows[0]; u = rows[1]; ix = {n: i for i, n in enumerate(h)}
W = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
     "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
out = ["# ncu --set full --clock-control none --import-source on, bench default workload (L-DGN, N=50, 32768 episodes/GPU, bf16, eager), one round; round 2, final kernels",
       "# (times under the profiler are cold-cache and serialised; the bench's event timings are the quoted numbers)", ""]
traffic = {}
mult = lambda unit: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
gemm_i = 0
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    out.append(f"== {name[:110]}  (grid {r[ix['launch__grid_size']]})")
    for w in W:
        if w in ix:
            out.append(f"  {w:100s} {r[ix[w]]:>16s} {u[ix[w]]}")
    out.append("")
    tot = float(r[ix["dram__bytes_read.sum"]]) * mult(u[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]]) * mult(u[ix["dram__bytes_write.sum"]])
    key = None
    if "conv2_attn" in name: key = "edge2"
    elif "attn_table_prep" in name: key = "edge1_prep"
    elif "attn_table" in name: key = "edge1"
    elif "env_round" in name: key = "env"
    elif "ctrl_need" in name: key = "ctrl_need_list"
    elif "gemm" in name:                       # epilogue mode 2 = first head layer, 3 = second (fused output), 1 = conv2 projections
        if "256, 2>" in name: key = "head0"
        elif "256, 3>" in name: key = "head1"
        else:
            key = ["proj2", "proj2_target"][gemm_i % 2]; gemm_i += 1
    if key: traffic[key] = int(round(tot, -3))
open("profiles/r02_ncu_kernels.txt", "w").write("\n".join(out))
json.dump({"model": "l_dgn", "episodes": 32768, "nodes": 50, "precision": "bf16",
           "source": "ncu --set full --clock-control none (one round of the eager bench, final kernels of round 2), dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "traffic": traffic}, open("profiles/r02_traffic.json", "w"), indent=1)
print(traffic)
