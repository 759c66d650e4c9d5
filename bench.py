#!/usr/bin/env python
"""Benchmark of the HGNN message-passing hot path (BASELINE.json metric:
"HGNN msg-passing edges/sec (fwd+bwd)").

Workload (BASELINE config 2): the isolated
gather -> edge-MLP -> scatter_add edge step, forward + backward, on a synthetic
symmetric random graph: E directed edges, N = E/10 nodes, latent L, hidden 2L,
2-layer LayerNorm/GELU/Tanh edge network (InteractionGNNCell.edge_update +
the scatter_add that feeds the next cell's node update). One "step" = one
fwd+bwd pass over all E edges; value = E * steps / time, summed over ranks
(each rank owns an independent batch of edges = data-parallel events; weight
gradients are all-reduced over NCCL, the only exchange the DP path has).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--latent L] [--edges E]
  python bench.py --impl reference      # CPU arm: the oracle port on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "hgnn_edge_step_fwd_bwd_edges_per_sec"
UNIT = "edge-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--edges", type=int, default=1_000_000)
    ap.add_argument("--impl", default="hgnn_b200", choices=["hgnn_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HGNN_PRECISION", "auto"))
    ap.add_argument("--cpu-edges", type=int, default=40_000, help="bounded CPU-baseline sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="pipelined", choices=["pipelined", "serial"],
                    help="pipelined: step i+1's H2D upload runs on a copy stream under step i's kernels; serial: same stream")
    ap.add_argument("--mode", default="dp", choices=["dp", "partition"],
                    help="dp: every rank owns a batch of edges (weak scaling, default); partition: ONE full-pile-up "
                         "event split by destination node across ranks (strong scaling, BASELINE config 5)")
    return ap.parse_args()


def hparams(L):
    return dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            p.update(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                     bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
        except Exception:  # noqa: BLE001
            pass
    return p


def workload_config(L, E, N, world):
    """The bench line's `config`, shared by both arms (the reference arm times a bounded sample of this workload)."""
    return {"workload": f"edge_step_fwd_bwd L={L} E={E} N=E/10 (BASELINE config 2)", "latent": L,
            "edges_per_gpu": E, "nodes_per_gpu": N, "l2_policy": "inputs (edge latents %d MB) larger than L2" % (E * L * 4 >> 20),
            "parallelism": f"dp{world}" if world > 1 else "single",
            "edge_order": "destination-sorted once per event (outside the step)"}


def alg_bytes_per_edge(L, n_over_e):
    """fwd 2*L*4 + 8 + 2*(N/E)*L*4 ; bwd 3*L*4 + 8 + 3*(N/E)*L*4."""
    fwd = 8 * L + 8 + 8 * L * n_over_e
    bwd = 12 * L + 8 + 12 * L * n_over_e
    return fwd, bwd


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference edge step on the host cores
# ---------------------------------------------------------------------------
def cpu_edge_step_rate(L, n_edges, reps, seed=42):
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from oracle import hgnn_oracle as O
    from oracle.reference_harness import kaiming_init
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = hparams(L)
    torch.manual_seed(0)
    from hierarchicalgnn_b200.utils import make_mlp
    net = make_mlp(3 * L, 2 * L, L, 2, layer_norm=True, output_activation="Tanh", hidden_activation="GELU")
    kaiming_init(net)
    sd = O.leaf_state({"edge_network." + k: v for k, v in net.state_dict().items()})
    nodes, edges, graph = synth_edge_problem(n_edges, L, seed=seed)
    times = []
    for i in range(reps + 1):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        O.edge_step_cell_fwd_bwd(sd, "edge_network", hp, nodes, edges, graph)
        times.append(time.perf_counter() - t0)
    times = times[1:]  # first call warms the allocator / MKL
    return n_edges / statistics.median(times), cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_edges
    per_step = []
    from hierarchicalgnn_b200.synth import synth_edge_problem  # noqa: F401
    rate, cores, times = cpu_edge_step_rate(args.latent, n, args.warmup + args.steps - 1)
    timed = times[-args.steps:] if len(times) >= args.steps else times
    t = sum(timed)
    value = n * len(timed) / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(timed), "warmup": args.warmup, "ms_per_step": 1e3 * t / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.latent, args.edges, max(2, int(round(args.edges * 0.1))), args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/hgnn_oracle.py edge step fwd+bwd on E={n} edges (N=E/10), L={args.latent}, "
                                   f"{len(timed)} timed passes, torch CPU fp32"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def run_gpu(args):
    import torch.distributed as dist
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from hierarchicalgnn_b200.training_utils import kaiming_init

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L, E = args.latent, args.edges
    hp = hparams(L)
    torch.manual_seed(0)
    cell = InteractionGNNCell(hp)
    kaiming_init(cell)
    cell.to(dev)
    net = cell.edge_network
    params = [p for p in net.parameters()]
    nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=42 + rank)
    N = nodes_h.shape[0]
    # edges are stored destination-sorted, as the models keep them (one sort per event, outside the cells): every
    # cell then streams edge rows in place and the segmented reduce sees contiguous runs
    order = torch.argsort(graph_h[1], stable=True)
    graph_h, edges_h = graph_h[:, order].contiguous(), edges_h[order].contiguous()
    g = torch.Generator().manual_seed(7 + rank)
    cot_e_h, cot_a_h = torch.randn(E, L, generator=g), torch.randn(N, L, generator=g)

    nodes = nodes_h.to(dev).requires_grad_(True)
    edges = edges_h.to(dev).requires_grad_(True)
    graph = graph_h.to(dev)
    cot_e, cot_a = cot_e_h.to(dev), cot_a_h.to(dev)
    gp = GraphPlans(graph, N, N)
    gp.by_src, gp.by_dst  # the segment plan is built once per graph, outside the step

    flat = None

    def step(nodes, edges, gp):
        """edge step fwd (+ the scatter_add feeding the next node update) and its backward."""
        e2, agg = net.edge_step(nodes, edges, gp.by_src, gp.by_dst)  # (e', agg) from one autograd node on the TC path
        if agg is None:
            agg = ops.scatter_add(e2, gp.graph[1], dim_size=N, plan=gp.by_dst)
        grads = torch.autograd.grad([e2, agg], [nodes, edges] + params, [cot_e, cot_a])
        if world > 1:
            nonlocal flat
            flat = torch.cat([x.reshape(-1) for x in grads[2:]])
            dist.all_reduce(flat)  # DP: weight gradients are the only exchanged data
        return e2, agg, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None  # started before warm-up so short timed regions are covered
    for _ in range(max(args.warmup, 3)):
        step(nodes, edges, gp)
    barrier()
    l0 = ops.LAUNCHES["count"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(nodes, edges, gp)
    ev1.record()
    barrier()
    launches = ops.LAUNCHES["count"] - l0
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * E / (ms_per_step * 1e-3)

    # ---- per-kernel timing of the step (CUDA events around each C-ABI call on the launch stream) ----
    roof = None
    ops.PROFILE = {}
    for _ in range(3):  # every rank runs these steps (they contain the gradient all-reduce); rank 0 reports
        step(nodes, edges, gp)
    torch.cuda.synchronize()
    prof = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in ops.PROFILE.items()}
    ops.PROFILE = None
    if rank == 0:
        pk = peaks()
        fwd_b, bwd_b = alg_bytes_per_edge(L, N / E)
        # dominant kernel of the step by measured time
        top = max(prof, key=prof.get) if prof else None
        # algorithmic FLOPs per edge: forward 16 L^2; the fp32 path recomputes the forward in its backward
        # (32 L^2 data + 16 L^2 weights), the tensor-core path stashes activations instead (16 L^2 data + 16 L^2 weights)
        flops_edge = {"mlp_forward": 16 * L * L, "mlp_backward_data": 32 * L * L, "mlp_backward_weights": 16 * L * L,
                      "tc_edge_forward": 16 * L * L, "tc_edge_backward": 32 * L * L}
        if top in flops_edge:
            ach = flops_edge[top] * E / (prof[top] * 1e-3) / 1e12
            peak = pk["bf16_tflops_sustained"]
            roof = {"bound": "tensor", "kernel": top, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": None, "peak_source": pk["source"] + " (cuBLAS bf16, sustained)"}
        elif top is not None:
            ach = (fwd_b + bwd_b) * E / (prof[top] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"]}
        if roof is not None:
            roof["kernel_ms"] = {k: round(v, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])}
            roof["step_hbm_frac"] = (fwd_b + bwd_b) * E / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"]
            step_flops = (48 if "tc_edge_backward" in prof else 64) * L * L  # 64 L^2 only where the backward recomputes
            roof["step_tensor_frac"] = step_flops * E / (ms_per_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]
            roof["step_flops_per_edge"] = step_flops
            # measured DRAM traffic of the dominant call from the committed ncu capture of this same workload
            tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
            if os.path.exists(tpath):
                try:
                    t = json.load(open(tpath)).get(top)
                    if t and t.get("edges") == E and t.get("latent") == L:
                        roof["traffic"] = t["dram_bytes_per_launch"]
                        roof["traffic_source"] = t.get("source")
                except Exception:  # noqa: BLE001
                    pass

    # ---- end to end through the public module API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.pin_memory()
        nodes_p, edges_p, graph_p, ce_p, ca_p = map(pin, (nodes_h, edges_h, graph_h, cot_e_h, cot_a_h))
        h2d = sum(t.numel() * t.element_size() for t in (nodes_p, edges_p, graph_p))
        out_host = torch.empty(2, dtype=torch.float32).pin_memory()

        copy_stream = torch.cuda.Stream(device=dev)
        # two preallocated device input sets (double buffer): uploads never allocate, so the step time does not depend on
        # what the caching allocator happens to hold
        slots = [tuple(torch.empty_like(t, device=dev) for t in (nodes_p, edges_p, graph_p)) for _ in range(2)]

        def upload(i):
            """H2D copy of step i's inputs from pinned host memory into device slot i % 2, on the copy stream (step i+1's
            upload runs under step i's kernels; every step's copy is inside the timed region). The slot's previous user,
            step i-2, has been synchronised by the host before this is issued."""
            bufs = slots[i % 2]
            with torch.cuda.stream(copy_stream):
                for dst_t, src_t in zip(bufs, (nodes_p, edges_p, graph_p)):
                    dst_t.copy_(src_t, non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy_stream)
            return bufs, done

        def e2e_step(cur):
            (n_b, e_b, g_d), done = cur
            torch.cuda.current_stream().wait_event(done)
            n_d = n_b.detach().requires_grad_(True)
            e_d = e_b.detach().requires_grad_(True)
            plans = GraphPlans(g_d, N, N)  # a new graph arrives with every event: plan build is inside the step
            e2, agg, grads = step(n_d, e_d, plans)
            # detached: out_host lives across steps, and an in-place copy of a tensor with history would chain every step's
            # autograd graph (and what its nodes hold) behind it
            metric = torch.stack([e2.detach().sum() + agg.detach().sum(), grads[1].abs().sum()])
            out_host.copy_(metric, non_blocking=True)

        if args.e2e_mode == "serial":
            copy_stream = torch.cuda.current_stream()

        def e2e_run(k):
            cur = upload(0)
            for i in range(k):
                nxt = upload(i + 1) if (i + 1 < k and args.e2e_mode == "pipelined") else None
                if args.e2e_mode == "serial" and i > 0:
                    cur = upload(i)
                e2e_step(cur)
                torch.cuda.current_stream().synchronize()  # the host reads the step's result before the next step
                cur = nxt
            return out_host

        e2e_run(3)
        barrier()
        k = max(3, min(args.steps, 10))
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_run(k)
        t1.record()
        barrier()
        ems = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * E / (ems / k * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
               "ms_per_step": ems / k, "steps": k,
               "pipeline": ("double-buffered H2D on a copy stream into preallocated device slots" if args.e2e_mode == "pipelined" else "H2D on the compute stream")
                           + ", host sync + 8 B D2H per step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, times = cpu_edge_step_rate(L, args.cpu_edges, 3)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle/hgnn_oracle.py edge step fwd+bwd, E={args.cpu_edges} edges (N=E/10), L={L}, median of 3, "
                         f"torch CPU fp32, {cores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ops.compute_dtype(net), "data": "synthetic",
            "config": workload_config(L, E, N, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_partition(args):
    """BASELINE config 5: one full-pile-up shaped event (N = 0.04 E), one InteractionGNNCell fwd+bwd, destination-
    partitioned over the ranks: all-gather of node rows forward, reduce-scatter of node gradients backward,
    all-reduce of weight gradients. Strong scaling: value = E_total / time."""
    import torch.distributed as dist
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.gnn_utils import InteractionGNNCell
    from hierarchicalgnn_b200.parallel import (allreduce_gradients, cuda_cell_callables, pad_rows, partition_by_destination,
                                               partitioned_interaction_cell)
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from hierarchicalgnn_b200.training_utils import kaiming_init
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = args.latent
    E = args.edges if args.edges != 1_000_000 else 3_000_000
    torch.manual_seed(0)
    cell = InteractionGNNCell(hparams(L))
    kaiming_init(cell)
    cell.to(dev)
    nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=2000, nodes_per_edge=0.04)
    N = nodes_h.shape[0]
    part = partition_by_destination(graph_h, N, world, rank)
    g = torch.Generator().manual_seed(11)
    cot_n, cot_e = torch.randn(N, L, generator=g), torch.randn(E, L, generator=g)
    own = slice(part.node_lo, part.node_hi)
    cot_n_d, cot_e_d = cot_n[own].to(dev), cot_e[part.edge_ids].to(dev)
    nodes = pad_rows(nodes_h, world * part.block).to(dev).requires_grad_(True)
    e_loc = edges_h[part.edge_ids].to(dev).requires_grad_(True)
    part.graph, part.dst_local, part.edge_ids = part.graph.to(dev), part.dst_local.to(dev), part.edge_ids.to(dev)
    node_fn, edge_fn, seg = cuda_cell_callables(cell)
    params = list(cell.parameters())

    def step():
        n2, e2 = partitioned_interaction_cell(part, nodes, e_loc, node_fn, edge_fn, seg)
        grads = torch.autograd.grad([n2[own], e2], [nodes, e_loc] + params, [cot_n_d, cot_e_d])
        if world > 1:
            flat = torch.cat([x.reshape(-1) for x in grads[2:]])
            dist.all_reduce(flat)
            # replicated input nodes: each rank holds the gradient of its own block only (the other rows are zero),
            # so the sum over ranks is an all-gather of the owned blocks
            gfull = torch.empty_like(grads[0])
            dist.all_gather_into_tensor(gfull, grads[0][part.node_lo:part.node_lo + part.block].contiguous())
        return n2, e2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.LAUNCHES["count"]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": E / (ms / args.steps * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": ops.compute_dtype(), "data": "synthetic",
            "config": {"workload": f"InteractionGNNCell fwd+bwd on one full-pile-up shaped event, L={L} E={E} N={N}, "
                                   f"destination-partitioned (BASELINE config 5)", "latent": L, "edges_total": E,
                       "nodes_total": N, "parallelism": f"dst-partition x{world}", "l2_policy": "inputs larger than L2",
                       "edges_rank0": int(part.edge_ids.numel())},
            "clocks": clocks, "e2e": None, "gpu_launches": ops.LAUNCHES["count"] - l0, "roofline": None, "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "partition":
        if args.precision != "auto":
            os.environ["HGNN_PRECISION"] = args.precision
        run_partition(args)
    else:
        if args.precision != "auto":
            os.environ["HGNN_PRECISION"] = args.precision
        run_gpu(args)


if __name__ == "__main__":
    main()
