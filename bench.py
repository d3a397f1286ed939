#!/usr/bin/env python
"""bench.py -- agent-transitions/s of the batched rollout round (env round + DGN forward +
epsilon-greedy action selection) on N B200s of one node.

    python bench.py [--gpus N --steps K --warmup W] [--model l_dgn|hl_dgn|dgn_r] [--impl reference]

A "step" is one round of the hot path over all resident episodes (BASELINE.json metric:
"agent-transitions/sec (env step + DGN forward), 50-node graphs").  Workload at N=1: the
per-GPU shard of BASELINE config 3 (L-DGN, 50-node graphs, 65536 episodes over 2 GPUs =
32768 episodes per GPU); more GPUs hold more episodes (weak scaling, no data-path collective).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "agent_transitions_per_sec"
UNIT = "agent-transitions/s"
CACHE_DIR = os.path.join(ROOT, "graph_topologies")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="l_dgn", choices=["l_dgn", "hl_dgn", "dgn_r", "none"])
    ap.add_argument("--nodes", type=int, default=50)
    ap.add_argument("--episodes", type=int, default=32768, help="episodes per GPU")
    ap.add_argument("--graphs", type=int, default=1024)
    ap.add_argument("--eps", type=float, default=0.05)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--preroll", type=int, default=32, help="untimed rounds so episodes are spread over their lifetime")
    ap.add_argument("--e2e-sub-batches", type=int, default=2, help="episode slices per host round (Rollout.round_host)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-episodes", type=int, default=8, help="episodes per CPU process in the CPU arm")
    ap.add_argument("--prof-kernel", default=None)
    ap.add_argument("--dynamic", action="store_true", help="dynamic_graph=True: nodes move every round (Philox stream on the device)")
    return ap.parse_args()


# ------------------------------------------------------------------------------ synthetic inputs
def load_pool(N, G):
    from melissa_b200.topology import GraphPool
    path = os.path.join(CACHE_DIR, f"pool_n{N}_{G}.npz")
    if os.path.exists(path):
        return GraphPool.load_npz(path)
    pool = GraphPool.synthetic(N, G, first_seed=0)
    try:
        os.makedirs(CACHE_DIR, exist_ok=True)
        pool.save_npz(path)
    except OSError:
        pass
    return pool


def _tuple_shard(args):
    from melissa_b200 import reset_chain
    base, count, N, G = args
    return reset_chain.episode_pool(base, count, N, G)


def load_tuples(N, G, count, base_seed=9, first_index=0):
    """Reset tuples drawn with the reference's RNG chain, env seed = base_seed + k (tianshou
    seeds vector env k with seed + k); graph (first_index + k) % G."""
    path = os.path.join(CACHE_DIR, f"tuples_n{N}_g{G}_c{count}_s{base_seed}.npz")
    if os.path.exists(path):
        z = np.load(path)
        return ((first_index + np.arange(count)) % G).astype(np.int32), z["src"], z["inter"], z["scr"]
    import multiprocessing as mp
    procs = max(1, min((os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1"))), 16))
    per = (count + procs - 1) // procs
    jobs = [(base_seed + p * per, min(per, count - p * per), N, G) for p in range(procs) if p * per < count]
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        parts = pool.map(_tuple_shard, jobs)
    gi = ((first_index + np.arange(count)) % G).astype(np.int32)
    src = np.concatenate([p[1] for p in parts])
    inter = np.concatenate([p[2] for p in parts])
    scr = np.concatenate([p[3] for p in parts])
    try:
        np.savez_compressed(path, gi=gi, src=src, inter=inter, scr=scr)
    except OSError:
        pass
    return gi, src, inter, scr


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        # samples that arrived inside the timed region; a region shorter than nvidia-smi's reporting period
        # falls back to every sample since the sampler started (the warm-up rounds are the same load)
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.02]
        window = "timed region"
        if not inside:
            inside, window = [r for _, r in self.rows], "warm-up + timed region (same load)"
        sm = [float(r[1]) for r in inside if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in inside if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------ our arm
def algorithmic_flops(model, N, E_plus_self, A):
    """SURVEY.md section 8d formulas, per graph-round."""
    if model == "hl_dgn":
        return N * 296192 + E_plus_self * 3072 + 328448
    if model == "l_dgn":
        return N * 1344768 + 2 * E_plus_self * 3072 + A * 656128
    if model == "dgn_r":
        return N * 2000128 + 2 * (E_plus_self - N) * 2048 + A * 656128
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    from melissa_b200 import _lib
    from melissa_b200.batched_env import BatchedGraphEnv, ResetTuplesDevice
    from melissa_b200.networks import NETWORKS
    from melissa_b200.rollout import Rollout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured" if peaks else "fallback"

    N, B, G = args.nodes, args.episodes, args.graphs
    pool = load_pool(N, G)
    P = 2 * B
    from melissa_b200.sharding import rank_seed_range, reduce_job
    # every rank draws its own pool: env seeds 9 + rank*P + k (global env index, tianshou seeds env i with seed + i),
    # graphs (rank*P + k) % G -- disjoint across ranks
    seed0, first = rank_seed_range(9, rank, P)
    gi, src, inter, scr = load_tuples(N, G, P, base_seed=seed0, first_index=first)
    env = BatchedGraphEnv(B, N, pool, device=dev, want_obs=True, dynamic_graph=args.dynamic)
    net = None
    if args.model != "none":
        torch.manual_seed(9)
        kw = dict(aggregator="max") if args.model == "hl_dgn" else {}
        net = NETWORKS[args.model](5, 128, 2, 4, N, dueling_param=({"hidden_sizes": [128, 128]}, {"hidden_sizes": [128, 128]}),
                                   device=str(dev), **kw).to(dev)
        net.set_precision(args.precision)
    ro = Rollout(env, net, eps=args.eps, seed=9 + rank)
    tuples_dev = ResetTuplesDevice(gi, src, inter, scr, N, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    use_graph = not args.no_graph

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def prepare(graph):
        """Identical starting point for every measured phase: same tuples, same Philox stream,
        `preroll` untimed rounds so the episodes are spread over their lifetime, then W warm-up rounds."""
        ro.graph = None
        ro.start(tuples_dev)
        if net is None:
            ro.act.copy_(torch.randint(0, 2, ro.act.shape, device=dev, dtype=torch.int8))
        n_pre = args.preroll - (2 if graph else 0)
        for _ in range(max(0, n_pre)):
            ro.round()
        if graph:
            ro.capture(warmup_rounds=2)
        for _ in range(args.warmup):
            ro.round()
        sync_all()

    two_convs = args.model in ("l_dgn", "dgn_r")
    # dominant kernel of the step: the conv2 attention pass (bf16, two-conv models), else the conv1 attention
    prof_name = args.prof_kernel or (("edge2" if two_convs else "edge1") if args.precision == "bf16" else ("proj2" if two_convs else "proj1"))
    # tensor-core line: conv2 source projection GEMM; HL-DGN (bf16, table mode) has no per-node projection GEMM
    # left, its largest GEMM is the first head layer
    gemm_name = "proj2" if two_convs else ("head0" if args.precision == "bf16" else "proj1")
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, with_prof=False, finish=None):
        """K steps, L2 flushed between steps (outside the per-step event pairs).  ``finish`` (optional) runs after
        the last step, inside the timed region: work the steps left in flight on other streams."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t_before = ro.transitions()
        l_before = _lib.lib().mls_launch_count()
        prof_ms, extra = [], []
        sync_all()
        for s in range(steps):
            flush.zero_()
            evs[s][0].record()
            extra.append(fn())
            evs[s][1].record()
            if with_prof:
                evs[s][1].synchronize()
                prof_ms.append(pe0.elapsed_time(pe1))
        ev_fin = None
        if finish is not None:
            finish()
            ev_fin = torch.cuda.Event(enable_timing=True)
            ev_fin.record()
        sync_all()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if ev_fin is not None:
            ms += evs[-1][1].elapsed_time(ev_fin)
        return ms, ro.transitions() - t_before, _lib.lib().mls_launch_count() - l_before, prof_ms, extra

    # ---- phase 1: the product path (CUDA-graph replay of the whole round unless --no-graph)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                      # before the pre-roll / warm-up rounds: nvidia-smi needs ~0.2 s to report
    prepare(use_graph)
    clocks.mark_begin()
    ms, trans, launches, _, _ = timed(ro.round, args.steps)
    clocks.mark_end()
    clk = clocks.stop() if rank == 0 else None
    if net is not None and ro.feature_violations() != 0:
        raise RuntimeError("discrete-feature mode: the environment produced non-integer feature columns")
    launches_per_step = None
    # ---- phase 2: same rounds launched eagerly with an event pair around one launch per step of
    #      (a) the conv2 attention kernel -- the single kernel with the largest share of the step --,
    #      (b) the conv2 source-side projection GEMM (the tensor-core kernel) and (c) the conv1 attention stage
    prof_ms, prof_ms_gemm, prof_ms_conv1, eager_ms = [], [], [], None
    if net is not None:
        prepare(False)
        net.set_profile_events(prof_name, pe0, pe1)
        eager_ms, _, launches_eager, prof_ms, _ = timed(ro.round, args.steps, with_prof=True)
        launches_per_step = launches_eager / args.steps
        if use_graph:
            launches = launches_eager          # a graph replay launches the same kernels; they are counted at capture
        if prof_name != gemm_name:
            prepare(False)
            net.set_profile_events(gemm_name, pe0, pe1)
            _, _, _, prof_ms_gemm, _ = timed(ro.round, args.steps, with_prof=True)
        else:
            prof_ms_gemm = prof_ms
        if args.precision == "bf16" and prof_name != "edge1":
            prepare(False)
            net.set_profile_events("edge1", pe0, pe1)
            _, _, _, prof_ms_conv1, _ = timed(ro.round, args.steps, with_prof=True)
        net.set_profile_events(None)
    # env kernel alone (HBM roofline of the environment round)
    env_evs = []
    for s in range(min(args.steps, 10)):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step_device(ro.act); b.record()
        env_evs.append((a, b))
    torch.cuda.synchronize()
    env_ms = float(np.mean([a.elapsed_time(b) for a, b in env_evs]))

    e2e = None
    if not args.no_e2e:
        prepare(False)
        ro.round_host(args.e2e_sub_batches)          # allocates the pinned mirrors
        if args.e2e_sub_batches > 1 and use_graph:
            ro.capture_host(args.e2e_sub_batches)    # one CUDA graph per episode slice
        prepare(False)
        ro.sync_host()
        # rounds are issued back to back: slice i of round k+1 waits only for slice i of round k to be back on the
        # host (the dependency of a caller feeding observations back); the timed region ends when the last copy lands
        pipelined = args.e2e_sub_batches > 1
        e_ms, e_trans, _, _, ex = timed(lambda: ro.round_host(args.e2e_sub_batches, wait=not pipelined), args.steps,
                                        finish=ro.host_drain if pipelined else None)
        e2e = (e_ms, e_trans, ex[0])

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    ms_max, trans_sum = reduce_job(ms, trans, dev)
    launches_sum = reduce(launches, dist.ReduceOp.SUM if world > 1 else None)
    if e2e is not None:
        e_ms_max, e_trans_sum = reduce_job(e2e[0], e2e[1], dev)

    out = None
    if rank == 0:
        steps = args.steps
        value = trans_sum / (ms_max / 1e3)
        A = trans / (steps * B)                                   # mean active agents per graph-round (this rank)
        deg = float(pool.adj.sum()) / len(pool)                   # directed edges per graph
        flops_round = algorithmic_flops(args.model, N, deg + N, A)
        # kernels bracketed with events (first chunk of every forward)
        HC, hid = 512, 128
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        # (profiles/r01_traffic.json), valid for the default workload only
        traffic = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("model") == args.model and tj.get("episodes") == B and tj.get("nodes") == N and tj.get("precision") == args.precision:
                traffic = tj.get("traffic", {})
        except (OSError, ValueError):
            pass
        import ctypes
        chunk_graphs = _lib.lib().mls_dgn_chunk_graphs(ctypes.byref(net._desc()), B) if net is not None else B
        chunk_rows = chunk_graphs * N
        nproj = 3 if args.model == "dgn_r" else 2
        # bf16: conv1 = all projections in one GEMM; conv2 = source-side projections on every node (the target
        # side runs on the controlling nodes only, in a second, smaller GEMM)
        n_out = HC if args.precision == "fp32" else (nproj * HC if gemm_name == "proj1" else (nproj - 1) * HC)
        gk = HC if gemm_name == "proj2" else hid
        gemm_rows = chunk_rows
        if gemm_name == "head0":                                    # [graphs x H*C] x [H*C x 2*128] (HL-DGN: one row per graph)
            gemm_rows, n_out, gk = chunk_graphs, 256, HC
        gemm_flops = 2.0 * gemm_rows * n_out * gk
        gemm_ms = float(np.mean(prof_ms_gemm)) if prof_ms_gemm else None
        roof_gemm = None
        if gemm_ms:
            ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
            roof_gemm = {"kernel": f"{args.precision} GEMM ({gemm_name}, [{gemm_rows}x{gk}]x[{gk}x{n_out}])",
                         "bound": "tensor", "achieved": round(ach, 3), "peak": tensor_peak, "unit": "TFLOP/s",
                         "frac": round(ach / tensor_peak, 5), "traffic": traffic.get(gemm_name),
                         "peak_source": f"{peak_src} (sustained bf16)", "kernel_ms": round(gemm_ms, 5), "launch_flops": gemm_flops}
        roofline = roof_gemm
        kern_ms = float(np.mean(prof_ms)) if prof_ms else None
        esz = 2
        ctrl_rows = A * chunk_graphs                                # controlling nodes (= agent transitions) per launch
        csr_bytes = chunk_graphs * (deg + 2 * (N + 1))              # neighbour lists + row pointers, once per graph

        def hbm_line(kernel, ms_k, nbytes, key, note):
            ach = nbytes / (ms_k * 1e-3) / 1e9
            return {"kernel": kernel, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(ach / hbm_peak, 4), "traffic": traffic.get(key), "peak_source": peak_src,
                    "kernel_ms": round(ms_k, 5), "launch_bytes": int(nbytes), "note": note}

        def conv1_bytes():
            # discrete-feature mode: the projections come from an L2-resident table; the kernel reads the keys
            # (4 B/node) and lists and writes relu(conv) (HC bf16 per node) + the controlling-node snapshot
            if args.model == "hl_dgn":                              # gather kernel + pooling: reads keys, writes one row per graph
                return chunk_rows * 4 + csr_bytes + chunk_graphs * HC * esz
            return chunk_rows * (4 + HC * esz) + csr_bytes + ctrl_rows * HC * esz

        roof_conv1 = None
        if kern_ms and prof_name == "edge2":
            # conv2 attention (edge_bf16_kernel, compact targets): reads the source-side projections of every node
            # ((nproj-1)*HC bf16), the target-side projection of the controlling nodes (HC bf16), the per-node dots,
            # slots and lists; writes the conv2 snapshot of the controlling nodes (HC bf16)
            nb = (chunk_rows * ((nproj - 1) * HC * esz + 4 * 4 + 4) + ctrl_rows * (HC * esz + 4 * 4 + 4) + csr_bytes +
                  ctrl_rows * HC * esz)
            roofline = hbm_line(f"edge_bf16_kernel (conv2 attention, {chunk_graphs} graphs x 4 heads, {int(ctrl_rows)} targets)", kern_ms, nb,
                                "edge2", "largest single-kernel share of the step; SIMT issue/latency bound (ncu: profiles/), not bandwidth bound")
            if prof_ms_conv1:
                c1_ms = float(np.mean(prof_ms_conv1))
                name = ("attn_table_mma_kernel + key compaction + pair-logit table (conv1 attention on tcgen05"
                        if N <= 62 and args.model != "hl_dgn" else "edge_bf16_kernel (conv1 attention, gather from the feature table")
                roof_conv1 = hbm_line(f"{name}, {chunk_graphs} graphs)", c1_ms, conv1_bytes(), "edge1",
                                      "event pair spans the whole conv1 attention stage")
        elif kern_ms and prof_name == "edge1":
            roofline = hbm_line(f"conv1 attention stage ({chunk_graphs} graphs x 4 heads)", kern_ms, conv1_bytes(), "edge1",
                                "largest share of the step")
        W = _lib.words_per_row(N)
        env_bytes = B * (N * (4 * W + 4 + 4 + 2 + 2 + 1 + 32 + 8 + 1 + 1) + 8 * 4 + 8 + 1)   # adj+pos(pool, L2) not counted
        env_roof = {"kernel": "env_round_kernel", "bound": "hbm", "traffic": traffic.get("env"), "achieved": round(env_bytes / (env_ms * 1e-3) / 1e9, 1),
                    "peak": hbm_peak, "unit": "GB/s", "frac": round(env_bytes / (env_ms * 1e-3) / 1e9 / hbm_peak, 4),
                    "kernel_ms": round(env_ms, 5), "launch_bytes": env_bytes, "peak_source": peak_src}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {
                "workload": f"{args.model} rollout round (env round + forward + eps-greedy), {N}-node graphs, "
                            f"{B} episodes per GPU (BASELINE config 3 per-GPU shard)",
                "episodes_per_gpu": B, "n_nodes": N, "dynamic_graph": bool(args.dynamic), "graph_pool": len(pool), "eps": args.eps, "precision": args.precision,
                "preroll_rounds": args.preroll, "l2": "flushed between timed steps (256 MiB memset outside the timed events)",
                "launch_mode": "cuda-graph replay of the whole round" if use_graph else "eager",
                "forward_mode": ("discrete-feature tables (MLS_FWD_DISCRETE_FEATURES)" if args.precision == "bf16" else "per-node"),
                "graphs_per_pass": int(chunk_graphs),
                "eager_ms_per_step": (eager_ms / steps) if eager_ms else None,
                "kernel_timing": "event pair around one launch per step in an eager pass over the same rounds",
                "active_agents_per_graph_round": round(A, 3), "graph_rounds_per_s": world * B * steps / (ms_max / 1e3),
                "model_tflops_algorithmic": round(flops_round * B * steps / (ms / 1e3) / 1e12, 3),
            },
            "gpu_launches": int(launches_sum),
            "clocks": clk,
            "roofline": roofline if roofline else env_roof,
            "roofline_tensor": roof_gemm,
            "roofline_conv1": roof_conv1,
            "roofline_env": env_roof,
        }
        if e2e is not None:
            out["e2e"] = {"value": e_trans_sum / (e_ms_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(e2e[2][0]),
                          "d2h_bytes_per_step": int(e2e[2][1]), "ms_per_step": e_ms_max / steps,
                          "pipeline": (f"{args.e2e_sub_batches} episode slices per round: H2D, compute and D2H of different slices overlap "
                                       "on three streams, slice i of round k+1 waits for slice i of round k to be back on the host "
                                       "(pinned host buffers; every byte of every round crosses PCIe inside the timed region, "
                                       "which ends when the last copy has landed)"
                                       if args.e2e_sub_batches > 1 else "one stream: H2D -> compute -> D2H")}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------ CPU arms
def cpu_arm(model, N, eps, episodes_per_proc, rounds, warmup):
    from oracle import cpu_rollout
    kind = None if model == "none" else model
    r = cpu_rollout.run_parallel(kind, N, episodes_per_proc, rounds, warmup, eps=eps)
    return r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    r = cpu_arm(args.model, args.nodes, args.eps, args.cpu_episodes, args.steps, args.warmup)
    sample = (f"oracle port (numpy env + per-agent-obs torch forward), {r['cores']} processes x {args.cpu_episodes} episodes, "
              f"{args.steps} rounds each after {args.warmup} warm-up rounds")
    return {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3 / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} rollout round, {args.nodes}-node graphs, CPU port of the reference path",
                   "n_nodes": args.nodes, "eps": args.eps},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    args = parse_args()
    if args.impl == "reference":
        out = run_reference(args)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    out = run_ours(args)
    if out is None:
        return
    if not args.no_cpu_baseline and args.gpus == 1:
        t0 = time.time()
        r = cpu_arm(args.model, args.nodes, args.eps, args.cpu_episodes, 4, 1)
        out["cpu_baseline"] = {
            "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"oracle port, {r['cores']} processes x {args.cpu_episodes} episodes x 4 rounds (1 warm-up), "
                      f"{r['transitions']} transitions in {r['seconds']:.1f}s (wall {time.time() - t0:.0f}s)"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
