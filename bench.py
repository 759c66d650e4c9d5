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

Besides `value` (device-timed, inputs resident) the line carries
  e2e            BASELINE config 3 through the public model API from HOST buffers: BC_HierarchicalGNN_GMM (latent 128,
                 6 + 6 cells) forward + backward on synthetic 1 GeV events, HGNN_BENCH_EVENTS_PER_STEP (default 16) events
                 collated into one disjoint graph per step (torch_geometric Batch layout), x[N,3] / edge_index / cluster
                 labels / event ids uploaded from pinned memory every step, the loss read back; same metric (edge-steps/s
                 = E_d * 12 cells / time). e2e.single_event = the same with ONE event per step (host-bound)
  models         device-timed and end-to-end times of BASELINE configs 1 (EC-IN forward) and 3 (BC fwd+bwd) on 1 GeV events
  dp_training    BASELINE config 4: BC training steps (loss, backward with the bucketed all-reduce overlapped,
                 clip 0.5, AdamW) on per-rank batches of 1 GeV events (HGNN_BENCH_EVENTS_PER_STEP, default 16, collated into
                 one disjoint graph; models.config3_bc_fwd_bwd_1gev_batched is the forward + backward of such a batch)
  partition      (N > 1) BASELINE config 5: one full-pile-up shaped event, destination-partitioned (strong scaling): a stack
                 of two InteractionGNNCells, and (partition.hierarchical) two HierarchicalGNNCells
  roofline, cpu_baseline   as the contract asks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--latent L] [--edges E]
  python bench.py --impl reference      # CPU arm: the reference's own InteractionGNNCell code on the host cores
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
    ap.add_argument("--cpu-edges", type=int, default=250_000, help="bounded sample size of the cpu_baseline leg of the GPU arm")
    ap.add_argument("--ref-edges", type=int, default=0, help="edges per step of --impl reference (0 = the full --edges workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-models", action="store_true", help="skip the config 1 / 3 / 4 / 5 sub-benchmarks")
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
# CPU arm. Nothing here imports the product package: inputs come from oracle/inputs.py, the timed code is the
# reference's own InteractionGNNCell (vendored by oracle/make_ref.py into oracle/_ref, kind "reference") or, where that
# copy is absent, the oracle restatement (kind "port").
# ---------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def have_vendored_reference():
    return os.path.exists(os.path.join(REF_DIR, "Modules", "gnn_utils.py"))


def _reference_classes():
    os.environ["HGNN_REFERENCE_ROOT"] = REF_DIR
    import importlib
    from oracle import reference_harness as rh
    importlib.reload(rh)  # picks the vendored root up even if the harness was imported earlier
    return rh, rh.reference_classes()


def cpu_edge_step_rate(L, n_edges, reps, kind, seed=42, warm=1):
    """Edge step + the scatter_add that feeds the next node update, forward + backward, fp32 on all host threads.
    kind "reference": gnn_utils.InteractionGNNCell.edge_update (checkpointed, as the reference runs it) + torch_scatter
    semantics from oracle/stubs; kind "port": oracle/hgnn_oracle.py. Returns (edge-steps/s from the median, cores, times)."""
    from oracle import hgnn_oracle as O
    from oracle.inputs import synth_edge_problem
    from oracle.seeded_state import seeded_init
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = hparams(L)
    nodes, edges, graph = synth_edge_problem(n_edges, L, seed=seed)
    N = nodes.shape[0]
    g = torch.Generator().manual_seed(7)
    cot_e, cot_a = torch.randn(n_edges, L, generator=g), torch.randn(N, L, generator=g)
    if kind == "reference":
        rh, C = _reference_classes()
        from torch_scatter import scatter_add  # oracle/stubs restatement (put on sys.path by the harness)
        cell = C["gnn_utils"].InteractionGNNCell(hp)
        seeded_init(cell, 0)

        def one():
            cell.zero_grad(set_to_none=True)
            n_, e_ = nodes.clone().requires_grad_(True), edges.clone().requires_grad_(True)
            e2 = cell.edge_update(n_, e_, graph)
            agg = scatter_add(e2, graph[1], dim=0, dim_size=N)
            ((e2 * cot_e).sum() + (agg * cot_a).sum()).backward()
    else:
        net = torch.nn.Sequential()  # parameter container with make_mlp's key layout (0,1,3,4)
        net.add_module("0", torch.nn.Linear(3 * L, 2 * L)); net.add_module("1", torch.nn.LayerNorm(2 * L))
        net.add_module("3", torch.nn.Linear(2 * L, L)); net.add_module("4", torch.nn.LayerNorm(L))
        seeded_init(net, 0)
        sd = O.leaf_state({"edge_network." + k: v for k, v in net.state_dict().items()})

        def one():
            for v in sd.values():
                v.grad = None
            n_, e_ = nodes.clone().requires_grad_(True), edges.clone().requires_grad_(True)
            e2 = O.edge_step(sd, "edge_network", hp, n_, e_, graph)
            agg = O.scatter_add(e2, graph[1], N)
            ((e2 * cot_e).sum() + (agg * cot_a).sum()).backward()
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    times = times[warm:]
    return n_edges / statistics.median(times), cores, times


def cpu_ec_forward_seconds(reps=2):
    """BASELINE config 1: the reference's EC_InteractionGNN.forward (latent 128, 14 cells) on one synthetic 1 GeV event,
    fp32, torch CPU with all host threads. Returns (best seconds, edge-steps per forward, cores) or None without oracle/_ref."""
    if not have_vendored_reference():
        return None
    from oracle.inputs import synth_event
    from oracle.seeded_state import seeded_init
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rh, C = _reference_classes()
    hp = rh.load_yaml_hparams("EC")
    model = C["EC_InteractionGNN"](hp)
    seeded_init(model, 0)
    model.eval()
    ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
    best = None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            model(ev.x.clone(), ev.edge_index)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return best, 2 * ev.edge_index.shape[1] * hp["n_interaction_graph_iters"], cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_edges or args.edges
    kind = "reference" if have_vendored_reference() else "port"
    rate, cores, times = cpu_edge_step_rate(args.latent, n, args.steps, kind, warm=max(1, min(args.warmup, 2)))
    t = sum(times)
    value = n * len(times) / t
    cfg = workload_config(args.latent, n, max(2, int(round(n * 0.1))), 1)
    cfg["parallelism"] = "host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": (f"{'reference gnn_utils.InteractionGNNCell.edge_update (checkpointed) + scatter_add' if kind == 'reference' else 'oracle/hgnn_oracle.py edge step + scatter_add'}"
                                    f", forward + backward, E={n} edges (N=E/10), L={args.latent}, {len(times)} timed passes, torch CPU fp32, {cores} threads")},
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


def _event_pool(n_events, seed0, n_particles=1200):
    """Pinned host copies of `n_events` synthetic 1 GeV events: x, edge_index, cluster labels (= particle)."""
    from hierarchicalgnn_b200.synth import synth_event
    pool = []
    for i in range(n_events):
        ev = synth_event(n_particles, 10, 0.0, 4.0, seed=seed0 + i)
        pool.append(dict(x=ev.x.pin_memory(), graph=ev.edge_index.pin_memory(), clusters=(ev.pid - 1).pin_memory(),
                         y=ev.y_pid.float().pin_memory(), pid=ev.pid.pin_memory(), pt=ev.pt.pin_memory(),
                         e_directed=2 * ev.edge_index.shape[1]))
    return pool


def _batch_pool(n_batches, events_per_batch, seed0, n_particles=1200):
    """Pinned host copies of `n_batches` collated batches of synthetic 1 GeV events (torch_geometric Batch layout:
    hierarchicalgnn_b200.synth.collate_events): x, edge_index, event id per hit, cluster labels (= particle), pid, pt."""
    from hierarchicalgnn_b200.synth import collate_events, synth_event
    pool = []
    for i in range(n_batches):
        evs = [synth_event(n_particles, 10, 0.0, 4.0, seed=seed0 + i * events_per_batch + j) for j in range(events_per_batch)]
        b = collate_events(evs)
        pool.append(dict(x=b.x.pin_memory(), graph=b.edge_index.pin_memory(), clusters=b.clusters.pin_memory(),
                         batch=b.batch.pin_memory(), pid=b.pid.pin_memory(), pt=b.pt.pin_memory(),
                         e_directed=2 * b.edge_index.shape[1], n_events=events_per_batch))
    return pool


def _events_per_step():
    return max(1, int(os.environ.get("HGNN_BENCH_EVENTS_PER_STEP", "16")))


def _timed(fn, k, barrier, world, dev):
    """k calls of fn(i) between two CUDA events on the current stream, barrier + synchronize on both sides; max over ranks."""
    import torch.distributed as dist
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(k):
        fn(i)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / k


def model_benchmarks(args, dev, world, rank, barrier):
    """BASELINE configs 1 and 3 on synthetic 1 GeV events (12 000 hits, ~108 k directed edges), each timed twice:
    device-timed with the event resident in HBM, and end to end from pinned HOST buffers with the result read back."""
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    out = {}
    pool = _event_pool(4, 1000 + 16 * rank)
    k = max(3, min(args.steps, 10))
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    # ---- config 3: BC_HierarchicalGNN_GMM, latent 128, forward + backward (supernodes = particles) ----
    torch.manual_seed(0)
    bc = model_selector("BC-HGNN-GMM", dict(latent=128))
    kaiming_init(bc)
    bc.to(dev).train()
    n_cells = bc.hparams["n_interaction_graph_iters"] + bc.hparams["n_hierarchical_graph_iters"]
    bc_params = [p for p in bc.parameters()]

    def bc_step(x, graph, clusters):
        for p in bc_params:
            p.grad = None
        bg, scores, emb = bc(x, graph, clusters=clusters)
        loss = scores.mean() + emb.square().mean()
        loss.backward()
        return loss.detach()

    resident = [tuple(ev[k_].to(dev) for k_ in ("x", "graph", "clusters")) for ev in pool]

    def bc_device(i):
        x, g, c = resident[i % len(resident)]
        bc_step(x.clone(), g, c)

    def bc_e2e(i):
        ev = pool[i % len(pool)]
        x, g, c = (ev[k_].to(dev, non_blocking=True) for k_ in ("x", "graph", "clusters"))
        loss_host.copy_(bc_step(x, g, c).reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the loss before the next event

    for i in range(3):
        bc_device(i)
        bc_e2e(i)
    dms = _timed(bc_device, k, barrier, world, dev)
    ems = _timed(bc_e2e, k, barrier, world, dev)
    es = sum(pool[i % len(pool)]["e_directed"] for i in range(k)) / k * n_cells
    h2d = sum(sum(pool[i % len(pool)][k_].numel() * pool[i % len(pool)][k_].element_size() for k_ in ("x", "graph", "clusters"))
              for i in range(k)) // k
    out["config3_bc_fwd_bwd_1gev"] = {
        "workload": f"BC_HierarchicalGNN_GMM latent 128, {n_cells} cells, forward + backward, synthetic 1 GeV events "
                    f"(12 000 hits, ~{int(es / n_cells)} directed edges; supernodes = particles), {world} GPU(s) x 1 event per step",
        "device_ms_per_step": dms, "e2e_ms_per_step": ems, "steps": k, "edge_steps_per_event": es,
        "device_edge_steps_per_s": world * es / (dms * 1e-3), "e2e_edge_steps_per_s": world * es / (ems * 1e-3),
        "events_per_s_e2e": world / (ems * 1e-3), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4}
    # ---- the same step on a collated batch of events (one disjoint graph per step: BASELINE config 4's "batches of events") ----
    B = _events_per_step()
    bpool = _batch_pool(2, B, 5000 + 64 * rank)
    keys = ("x", "graph", "clusters", "batch")

    def bc_batch_step(x, graph, clusters, batch):
        for p in bc_params:
            p.grad = None
        bg, scores, emb = bc(x, graph, clusters=clusters, batch=batch, n_events=B)
        loss = scores.mean() + emb.square().mean()
        loss.backward()
        return loss.detach()

    bres = [tuple(b[k_].to(dev) for k_ in keys) for b in bpool]

    def bcb_device(i):
        x, g, c, bt = bres[i % len(bres)]
        bc_batch_step(x.clone(), g, c, bt)

    def bcb_e2e(i):
        b = bpool[i % len(bpool)]
        x, g, c, bt = (b[k_].to(dev, non_blocking=True) for k_ in keys)
        loss_host.copy_(bc_batch_step(x, g, c, bt).reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(3):
        bcb_device(i)
        bcb_e2e(i)
    dms = _timed(bcb_device, k, barrier, world, dev)
    ems = _timed(bcb_e2e, k, barrier, world, dev)
    es = sum(bpool[i % len(bpool)]["e_directed"] for i in range(k)) / k * n_cells
    h2d = sum(sum(bpool[i % len(bpool)][k_].numel() * bpool[i % len(bpool)][k_].element_size() for k_ in keys) for i in range(k)) // k
    out["config3_bc_fwd_bwd_1gev_batched"] = {
        "workload": f"BC_HierarchicalGNN_GMM latent 128, {n_cells} cells, forward + backward, {B} synthetic 1 GeV events collated "
                    f"into one disjoint graph per step (supernodes = particles; kNN graphs searched per event), {world} GPU(s)",
        "events_per_step_per_gpu": B, "device_ms_per_step": dms, "e2e_ms_per_step": ems, "steps": k, "edge_steps_per_step": es,
        "device_edge_steps_per_s": world * es / (dms * 1e-3), "e2e_edge_steps_per_s": world * es / (ems * 1e-3),
        "events_per_s_e2e": world * B / (ems * 1e-3), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4}
    del bc, bc_params, resident, bres

    # ---- config 1: EC_InteractionGNN forward (inference), latent 128, 14 cells ----
    if not args.no_models:
        torch.manual_seed(0)
        ec = model_selector("EC-IN")
        kaiming_init(ec)
        ec.to(dev).eval()
        n_cells = ec.hparams["n_interaction_graph_iters"]
        resident = [tuple(ev[k_].to(dev) for k_ in ("x", "graph")) for ev in pool]
        score_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def ec_device(i):
            x, g = resident[i % len(resident)]
            with torch.no_grad():
                ec(x, g)

        def ec_e2e(i):
            ev = pool[i % len(pool)]
            x, g = ev["x"].to(dev, non_blocking=True), ev["graph"].to(dev, non_blocking=True)
            with torch.no_grad():
                score_host.copy_(ec(x, g).mean().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for i in range(3):
            ec_device(i)
            ec_e2e(i)
        dms = _timed(ec_device, k, barrier, world, dev)
        ems = _timed(ec_e2e, k, barrier, world, dev)
        es = sum(pool[i % len(pool)]["e_directed"] for i in range(k)) / k * n_cells
        out["config1_ec_forward_1gev"] = {
            "workload": f"EC_InteractionGNN latent 128, {n_cells} cells, forward only, synthetic 1 GeV events",
            "device_ms_per_step": dms, "e2e_ms_per_step": ems, "steps": k, "edge_steps_per_event": es,
            "device_edge_steps_per_s": world * es / (dms * 1e-3), "e2e_edge_steps_per_s": world * es / (ems * 1e-3)}
        # the same forward on a collated batch (a flat interaction network needs no event vector: the graph is disjoint)
        bres = [(b["x"].to(dev), b["graph"].to(dev)) for b in bpool]

        def ecb_device(i):
            x, g = bres[i % len(bres)]
            with torch.no_grad():
                ec(x, g)

        def ecb_e2e(i):
            b = bpool[i % len(bpool)]
            x, g = b["x"].to(dev, non_blocking=True), b["graph"].to(dev, non_blocking=True)
            with torch.no_grad():
                score_host.copy_(ec(x, g).mean().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for i in range(3):
            ecb_device(i)
            ecb_e2e(i)
        dms = _timed(ecb_device, k, barrier, world, dev)
        ems = _timed(ecb_e2e, k, barrier, world, dev)
        es = sum(bpool[i % len(bpool)]["e_directed"] for i in range(k)) / k * n_cells
        out["config1_ec_forward_1gev_batched"] = {
            "workload": f"EC_InteractionGNN latent 128, {n_cells} cells, forward only, {B} synthetic 1 GeV events collated per step",
            "events_per_step_per_gpu": B, "device_ms_per_step": dms, "e2e_ms_per_step": ems, "steps": k, "edge_steps_per_step": es,
            "device_edge_steps_per_s": world * es / (dms * 1e-3), "e2e_edge_steps_per_s": world * es / (ems * 1e-3),
            "events_per_s_e2e": world * B / (ems * 1e-3)}
    return out


def dp_training_benchmark(args, dev, world, rank, barrier):
    """BASELINE config 4: HGNN_GMM (latent 128) TRAINING on per-rank synthetic 1 GeV events, data-parallel: training_step
    (hinge embedding loss + assignment BCE with the scipy matching on the host, as the reference), backward with the
    bucketed gradient all-reduce launched from hooks, global-norm clip 0.5 (Notebooks/script.py:35), AdamW step, buffer
    broadcast. Supernodes = particles (clustering injected)."""
    from types import SimpleNamespace
    from hierarchicalgnn_b200.parallel import DataParallelTrainer
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    torch.manual_seed(0)
    model = model_selector("BC-HGNN-GMM", dict(latent=128, loss_schedule=0.5))
    kaiming_init(model)
    model.to(dev).train()
    opt = model.configure_optimizers()[0][0]
    trainer = DataParallelTrainer(model, opt, clip=0.5, bucket_bytes=4 << 20)
    B = _events_per_step()
    pool = _batch_pool(3, B, 3000 + 64 * rank)
    dev_pool = []
    for b in pool:
        dev_pool.append(SimpleNamespace(x=b["x"].to(dev), edge_index=b["graph"].to(dev), pid=b["pid"].to(dev), pt=b["pt"].to(dev),
                                        clusters=b["clusters"].to(dev), batch=b["batch"].to(dev), num_graphs=B))
    cur = {}
    model.hgnn_block.clustering = lambda x, emb, graph: cur["clusters"]  # supernodes = particles

    def step(i):
        b = dev_pool[i % len(dev_pool)]
        cur["clusters"] = b.clusters
        b.x = b.x.detach().clone()
        trainer.step(b)

    for i in range(3):
        step(i)
    k = max(3, min(args.steps, 10))
    ms = _timed(step, k, barrier, world, dev)
    n_cells = model.hparams["n_interaction_graph_iters"] + model.hparams["n_hierarchical_graph_iters"]
    es = sum(pool[i % len(pool)]["e_directed"] for i in range(k)) / k * n_cells
    grad_bytes = sum(b["flat"].numel() for b in trainer.buckets.buckets) * 4
    res = {"workload": f"BC_HierarchicalGNN_GMM latent 128 training step (loss + backward + overlapped all-reduce + clip 0.5 + AdamW), "
                       f"{B} synthetic 1 GeV events per GPU per step (collated into one disjoint graph), {world} GPU(s)",
           "events_per_step_per_gpu": B,
           "ms_per_step": ms, "steps": k, "events_per_s": world * B / (ms * 1e-3), "edge_steps_per_s": world * es / (ms * 1e-3),
           "gradient_bytes": grad_bytes, "buckets": len(trainer.buckets.buckets),
           "timing": "wall of the training loop on the device clock (CUDA events, max over ranks); includes the host-side scipy matching"}
    trainer.buckets.remove()
    return res


def partition_benchmark(args, dev, world, rank, barrier):
    """BASELINE config 5 inside the default line: one full-pile-up shaped event (E = 3 M directed edges, N = 120 k), one
    interaction cell forward + backward, destination-partitioned over the ranks; the one-GPU time of the same event is
    measured in the same run (every rank runs it redundantly) so that the speed-up is self-contained."""
    res = run_partition(args, dev=dev, world=world, rank=rank, barrier=barrier, emit=False, steps=max(3, min(args.steps, 10)))
    if os.environ.get("HGNN_PART_HIER", "1") != "0":
        hier = hier_partition_benchmark(dev, world, rank, barrier, steps=max(3, min(args.steps, 6)))
        if res is not None:
            res["hierarchical"] = hier
    return res


def hier_partition_benchmark(dev, world, rank, barrier, steps=5, E=3_000_000, L=128):
    """The same strong-scaling measurement for the WHOLE hierarchical cell (gnn_utils.py:112-157): two HierarchicalGNNCells
    forward + backward on the full-pile-up shaped event (hits and hit edges destination-partitioned, the supernode side —
    S = N / 10 supernodes, N superedges, 3 hit -> supernode assignments per hit — replicated, its messages all-reduced),
    against the un-partitioned cells on one GPU in the same run. Parity of the two is `profiles/hier_partition_check.py`."""
    import torch.distributed as dist
    from hierarchicalgnn_b200.gnn_utils import GraphPlans, HierarchicalGNNCell
    from hierarchicalgnn_b200.parallel import (SymmetricRows, cuda_hier_cell_callables, pad_rows, partition_bipartite,
                                               partition_by_destination, partitioned_hierarchical_cell)
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from hierarchicalgnn_b200.training_utils import kaiming_init
    torch.manual_seed(0)
    cells = [HierarchicalGNNCell(hparams(L)) for _ in range(2)]
    for c in cells:
        kaiming_init(c)
        c.to(dev)
    nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=2000, nodes_per_edge=0.04)
    N = nodes_h.shape[0]
    g = torch.Generator().manual_seed(5)
    S, ES = max(N // 10, 8), max(N, 64)
    sn_h, se_h = torch.randn(S, L, generator=g), torch.randn(ES, L, generator=g)
    sg_h, sw_h = torch.randint(0, S, (2, ES), generator=g), torch.rand(ES, 1, generator=g)
    bg_h = torch.stack([torch.arange(N).repeat(3), torch.randint(0, S, (3 * N,), generator=g)])
    bw_h = torch.rand(3 * N, 1, generator=g) / 3
    cots = [torch.randn(*sh, generator=g) for sh in ((N, L), (E, L), (S, L), (ES, L))]
    order = torch.argsort(graph_h[1], stable=True)
    graph_h, edges_h, cots[1] = graph_h[:, order].contiguous(), edges_h[order].contiguous(), cots[1][order].contiguous()

    def leaf(t):
        return t.to(dev).clone().requires_grad_(True)
    params = [p for c in cells for p in c.parameters()]
    sg, cd = sg_h.to(dev), [c.to(dev) for c in cots]

    # ---- un-partitioned, every rank redundantly ----
    r = dict(n=leaf(nodes_h), e=leaf(edges_h), s=leaf(sn_h), se=leaf(se_h), bw=leaf(bw_h), sw=leaf(sw_h))
    graph, bg = graph_h.to(dev), bg_h.to(dev)
    gp, bp, sp = GraphPlans(graph, N, N, dst_sorted=True), GraphPlans(bg, N, S), GraphPlans(sg, S, S)

    def solo(_i=0):
        a, b, c, d = r["n"], r["e"], r["s"], r["se"]
        for i, cell in enumerate(cells):
            a, b, c, d = cell(a, b, c, d, gp, bp, r["bw"], sp, r["sw"], skip_edge_updates=(i == 1))
        torch.autograd.grad(sum((o * ct).sum() for o, ct in zip((a, b, c, d), cd)), list(r.values()) + params, allow_unused=True)
    for i in range(2):
        solo()
    t1 = _timed(solo, steps, barrier, world, dev)
    del r, graph, bg, gp, bp, sp

    # ---- partitioned ----
    part = partition_by_destination(graph_h, N, world, rank)
    bpart = partition_bipartite(bg_h, part)
    own = slice(part.node_lo, part.node_hi)
    p = dict(n=leaf(pad_rows(nodes_h, world * part.block)), e=leaf(edges_h[part.edge_ids]), s=leaf(sn_h), se=leaf(se_h),
             bw=leaf(bw_h[bpart.ids]), sw=leaf(sw_h))
    cn, ce = cd[0][own], cd[1][part.edge_ids.to(dev)]
    for k_ in ("graph", "dst_local", "edge_ids"):
        setattr(part, k_, getattr(part, k_).to(dev))
    bpart.ids, bpart.node_local, bpart.supernode = bpart.ids.to(dev), bpart.node_local.to(dev), bpart.supernode.to(dev)
    fns = [cuda_hier_cell_callables(c, fuse_aggregate=(i == 0)) for i, c in enumerate(cells)]
    sr = SymmetricRows(part.block, L, dev, slots=2) if world > 1 else None

    def parted(_i=0):
        st = dict(nodes=p["n"], edges=p["e"], supernodes=p["s"], superedges=p["se"], agg_owned=None, x_owned=None)
        for i, f in enumerate(fns):
            st = partitioned_hierarchical_cell(part, bpart, st["nodes"], st["edges"], st["supernodes"], st["superedges"], p["bw"],
                                               sg, p["sw"], f, symmetric=sr, agg_owned=st["agg_owned"], x_owned=st["x_owned"],
                                               slot=i, skip_edge_updates=(i == 1))
        loss = ((st["nodes"][own] * cn).sum() + (st["edges"] * ce).sum() + (st["supernodes"] * cd[2]).sum()
                + (st["superedges"] * cd[3]).sum())
        grads = torch.autograd.grad(loss, list(p.values()) + params, allow_unused=True)
        flat = torch.cat([x.reshape(-1) for x in grads[6:] if x is not None])
        if world > 1:
            dist.all_reduce(flat)  # the step-end weight-gradient all-reduce
    for i in range(2):
        parted()
    tp = _timed(parted, steps, barrier, world, dev)
    return {"workload": f"2 HierarchicalGNNCells fwd+bwd, L={L} E={E} N={N} S={S} supernodes, {ES} superedges, {3 * N} hit-supernode "
                        f"assignments; hits and hit edges destination-partitioned x{world}, supernode side replicated",
            "one_gpu_ms_per_step": t1, "ms_per_step": tp, "speedup_vs_1gpu": t1 / tp, "steps": steps}


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
            tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if not os.path.exists(tpath):
                tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
            if os.path.exists(tpath):
                try:
                    t = json.load(open(tpath)).get(top)
                    if t and t.get("edges") == E and t.get("latent") == L:
                        roof["traffic"] = t["dram_bytes_per_launch"]
                        roof["traffic_source"] = t.get("source")
                except Exception:  # noqa: BLE001
                    pass

    # ---- end to end through the public MODEL API with HOST buffers: BASELINE config 3 (and config 1) on 1 GeV events ----
    e2e, models = None, None
    if not args.no_e2e:
        models = model_benchmarks(args, dev, world, rank, barrier)
        m3, m1 = models["config3_bc_fwd_bwd_1gev_batched"], models["config3_bc_fwd_bwd_1gev"]
        e2e = {"value": m3["e2e_edge_steps_per_s"], "unit": UNIT, "h2d_bytes_per_step": m3["h2d_bytes_per_step"],
               "d2h_bytes_per_step": m3["d2h_bytes_per_step"], "ms_per_step": m3["e2e_ms_per_step"], "steps": m3["steps"],
               "device_ms_per_step": m3["device_ms_per_step"], "events_per_step_per_gpu": m3["events_per_step_per_gpu"],
               "workload": m3["workload"],
               "pipeline": "per step: x[N,3] f32 + edge_index[2,E] i64 + cluster labels[N] i64 + event id per hit[N] i64 of the collated "
                           "events copied from pinned host memory on the compute stream, model forward + backward (model(x, edge_index, "
                           "clusters=, batch=)), 4 B loss read back (host sync)",
               "single_event": {"value": m1["e2e_edge_steps_per_s"], "ms_per_step": m1["e2e_ms_per_step"],
                                "device_ms_per_step": m1["device_ms_per_step"], "h2d_bytes_per_step": m1["h2d_bytes_per_step"],
                                "workload": m1["workload"]}}
    extra = {}
    if not args.no_models:
        extra["dp_training"] = dp_training_benchmark(args, dev, world, rank, barrier)
    if world > 1 and not args.no_models:
        extra["partition"] = partition_benchmark(args, dev, world, rank, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = "reference" if have_vendored_reference() else "port"
        rate, cores, times = cpu_edge_step_rate(L, args.cpu_edges, 5, kind)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{'reference gnn_utils.InteractionGNNCell.edge_update (checkpointed) + scatter_add' if kind == 'reference' else 'oracle port'}"
                         f", forward + backward, E={args.cpu_edges} edges (N=E/10), L={L}, 1 warm-up + median of 5, torch CPU fp32, {cores} threads"}
        if kind == "reference":  # the restatement timed beside the reference's own code: validates the port's speed
            prate, _, _ = cpu_edge_step_rate(L, args.cpu_edges, 3, "port")
            cpu["port_value"] = prate
        ec = cpu_ec_forward_seconds(2) if not args.no_models else None
        if ec is not None:
            cpu["config1_ec_forward_1gev"] = {"seconds": ec[0], "edge_steps_per_s": ec[1] / ec[0], "cores": ec[2],
                                              "what": "reference EC_InteractionGNN.forward (latent 128, 14 cells), one synthetic 1 GeV event, fp32, best of 2"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ops.compute_dtype(net), "data": "synthetic",
            "config": workload_config(L, E, N, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "models": models,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_partition(args, dev=None, world=None, rank=None, barrier=None, emit=True, steps=None):
    """BASELINE config 5: one full-pile-up shaped event (N = 0.04 E), one InteractionGNNCell fwd+bwd, destination-
    partitioned over the ranks: all-gather of node rows forward, reduce-scatter of node gradients backward,
    all-reduce of weight gradients. Strong scaling: value = E_total / time. The same event is also run un-partitioned on
    one GPU (every rank, redundantly, in a group of its own) so that the record carries its own 1-GPU denominator."""
    import torch.distributed as dist
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.gnn_utils import InteractionGNNCell
    from hierarchicalgnn_b200.parallel import (SymmetricRows, cuda_cell_callables, pad_rows, partition_by_destination,
                                               partitioned_interaction_cell)
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from hierarchicalgnn_b200.training_utils import kaiming_init
    own_pg = dev is None
    if own_pg:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
    steps = steps or args.steps
    L = args.latent
    E = args.edges if (own_pg and args.edges != 1_000_000) else 3_000_000
    n_cells = max(1, int(os.environ.get("HGNN_PART_CELLS", "2")))
    torch.manual_seed(0)
    cells = []
    for _ in range(n_cells):
        c = InteractionGNNCell(hparams(L))
        kaiming_init(c)
        cells.append(c.to(dev))
    nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=2000, nodes_per_edge=0.04)
    N = nodes_h.shape[0]
    g = torch.Generator().manual_seed(11)
    cot_n, cot_e = torch.randn(N, L, generator=g), torch.randn(E, L, generator=g)
    # every cell but the last leaves scatter_add(e', dst) behind for the next cell's node update (fused segmented reduce)
    calls = [cuda_cell_callables(c, fuse_aggregate=(i + 1 < n_cells)) for i, c in enumerate(cells)]
    params = [p for c in cells for p in c.parameters()]
    solo = None
    if world > 1:
        groups = [dist.new_group([r]) for r in range(world)]  # every rank creates every group (collective call)
        solo = groups[rank]
    use_symm = world > 1 and os.environ.get("HGNN_PART_SYMM", "1") != "0"
    coll = {"impl": "nccl"}

    def make_step(w, r, group):
        part = partition_by_destination(graph_h, N, w, r)
        own = slice(part.node_lo, part.node_hi)
        cot_n_d, cot_e_d = cot_n[own].to(dev), cot_e[part.edge_ids].to(dev)
        nodes = pad_rows(nodes_h, w * part.block).to(dev).requires_grad_(True)
        e_loc = edges_h[part.edge_ids].to(dev).requires_grad_(True)
        part.graph, part.dst_local, part.edge_ids = part.graph.to(dev), part.dst_local.to(dev), part.edge_ids.to(dev)
        sr = None
        if w > 1 and use_symm:
            try:  # the repo's own peer-memory row collectives (csrc/p2p.cu) over symmetric tables; NCCL otherwise
                sr = SymmetricRows(part.block, L, dev, group=group, slots=n_cells)
                coll["impl"] = "peer-memory kernels, multimem (NVLS)" if sr.mc_base else "peer-memory kernels, peer pointers"
            except Exception as ex:  # noqa: BLE001
                coll["impl"] = f"nccl (symmetric memory unavailable: {type(ex).__name__})"
                sr = None

        def step(_i=0):
            x, e, agg, xo = nodes, e_loc, None, None
            for ci, (node_fn, edge_fn, seg) in enumerate(calls):
                x, e, agg, xo = partitioned_interaction_cell(part, x, e, node_fn, edge_fn, seg, group=group, symmetric=sr,
                                                             agg_owned=agg, return_agg=True, x_owned=xo, slot=ci,
                                                             return_owned=True)
            grads = torch.autograd.grad([x[own], e], [nodes, e_loc] + params, [cot_n_d, cot_e_d])
            if w > 1:
                flat = torch.cat([t.reshape(-1) for t in grads[2:]])
                if sr is not None:
                    sr.all_reduce_(flat)
                else:
                    dist.all_reduce(flat, group=group)
                # replicated input nodes: each rank holds the gradient of its own block only (the other rows are zero),
                # so the sum over ranks is an all-gather of the owned blocks
                blk = grads[0][part.node_lo:part.node_lo + part.block].contiguous()
                if sr is not None:
                    sr.all_gather(blk)
                else:
                    gfull = torch.empty_like(grads[0])
                    dist.all_gather_into_tensor(gfull, blk, group=group)
            return x, e
        return step, part, sr

    one_gpu_ms = None
    if world > 1:  # the 1-GPU denominator, measured here and now with the same kernels
        step1, _, _ = make_step(1, 0, solo)
        for _ in range(3):
            step1()
        one_gpu_ms = _timed(step1, steps, barrier, world, dev)
        del step1
        torch.cuda.empty_cache()
    step, part, sr = make_step(world, rank, None)
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(dev.index) if (rank == 0 and own_pg) else None
    l0 = ops.LAUNCHES["count"]
    ms = _timed(step, steps, barrier, world, dev)
    launches = ops.LAUNCHES["count"] - l0
    clocks = sampler.stop() if sampler else None
    # the two row collectives of a cell on their own (same tables, same barriers), for the record
    if world > 1:
        blk = torch.randn(part.block, L, device=dev)
        full = torch.randn(world * part.block, L, device=dev)
        if sr is not None:
            coll["all_gather_ms"] = _timed(lambda i: sr.all_gather(blk), 10, barrier, world, dev)
            coll["reduce_scatter_ms"] = _timed(lambda i: sr.reduce_scatter(full), 10, barrier, world, dev)
        else:
            o1, o2 = torch.empty_like(full), torch.empty_like(blk)
            coll["all_gather_ms"] = _timed(lambda i: dist.all_gather_into_tensor(o1, blk), 10, barrier, world, dev)
            coll["reduce_scatter_ms"] = _timed(lambda i: dist.reduce_scatter_tensor(o2, full), 10, barrier, world, dev)
        coll["table_bytes"] = int(world * part.block * L * 4)
    res = {"workload": f"{n_cells} InteractionGNNCell(s) fwd+bwd on one full-pile-up shaped event, L={L} E={E} N={N}, destination-"
                       f"partitioned x{world} (BASELINE config 5)", "ms_per_step": ms, "steps": steps,
           "cells": n_cells, "edge_steps_per_s": n_cells * E / (ms * 1e-3),
           "one_gpu_ms_per_step": one_gpu_ms, "speedup_vs_1gpu": (one_gpu_ms / ms) if one_gpu_ms else None,
           "edges_rank0": int(part.edge_ids.numel()), "scaling": "strong", "collectives": coll if world > 1 else None}
    if emit and rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": n_cells * E / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": ops.compute_dtype(), "data": "synthetic",
            "config": {"workload": res["workload"], "latent": L, "edges_total": E,
                       "nodes_total": N, "parallelism": f"dst-partition x{world}", "l2_policy": "inputs larger than L2",
                       "edges_rank0": res["edges_rank0"]},
            "clocks": clocks, "e2e": None, "gpu_launches": launches, "roofline": None, "cpu_baseline": None,
            "partition": res}))
    if own_pg and world > 1:
        dist.destroy_process_group()
    return res


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
