"""Microbenchmarks of every kernel family on the hot path, one B200, CUDA events, 3 warm-up + 10 timed calls,
inputs larger than L2 where the size allows. Prints a markdown table: achieved GB/s on the algorithmic bytes for the
HBM-bound kernels (against MEASURED_PEAKS.json), edge-steps/s for the fused edge step across latents / sizes / a
power-law graph, pairs/s for the brute-force kNN.  Usage: python profiles/microbench.py > profiles/r01_microbench.md"""
import json, os, sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem, synth_event, direction_embeddings
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector

PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6554.2
DEV = 'cuda'


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("# Microbenchmarks (one B200; `python profiles/microbench.py`)\n")
print(f"HBM peak used for fractions: {PEAK:.0f} GB/s (MEASURED_PEAKS.json)\n")
print("## HBM-bound kernels (algorithmic bytes / time)\n")
print("| kernel | shape | ms | GB/s | of peak |\n|---|---|---:|---:|---:|")
g = torch.Generator().manual_seed(0)
for E, N, L in [(4_000_000, 400_000, 128), (1_000_000, 100_000, 128), (120_000, 12_000, 128)]:
    src = torch.randn(E, L, generator=g).to(DEV)
    idx = torch.randint(0, N, (E,), generator=g).to(DEV)
    idx_sorted = idx.sort().values
    for name, keys in (("random segment ids", idx), ("sorted segment ids", idx_sorted)):
        plan = ops.plan_for(keys, N)
        t = timeit(lambda: ops.segment_reduce_raw(src, plan))
        by = E * L * 4 + N * L * 4 + E * 4
        print(f"| `hgnn_segment_reduce` ({name}) | E={E:,} N={N:,} L={L} | {t:.3f} | {by / t / 1e6:.0f} | {by / t / 1e6 / PEAK:.0%} |")
    x = torch.randn(N, L, generator=g).to(DEV)
    plan = ops.plan_for(idx, N)
    t = timeit(lambda: ops.gather_rows_raw(x, plan.keys32, None, E))
    by = E * L * 4 + E * 4 + N * L * 4
    print(f"| `hgnn_gather_rows` | E={E:,} N={N:,} L={L} | {t:.3f} | {by / t / 1e6:.0f} | {by / t / 1e6 / PEAK:.0%} |")
    t = timeit(lambda: ops.edge_dot_raw(x, plan.keys32, x, plan.keys32, E))
    by = E * 8 + E * 4  # node rows are L2-resident: indices in, one float out
    print(f"| `hgnn_edge_dot` (rows from L2) | E={E:,} N={N:,} L={L} | {t:.3f} | {2 * E * L * 4 / t / 1e6:.0f} (gathered) | — |")
    t = timeit(lambda: ops.SegmentPlan(idx, N))
    print(f"| `hgnn_csr_build` | {E:,} int64 keys | {t:.3f} | {E / t / 1e3:.0f} M keys/s | — |")
    del src, x

print("\n## Graph construction\n")
print("| kernel | shape | ms | rate |\n|---|---|---:|---:|")
for nq, nr, k in [(12_000, 1_200, 5), (1_200, 1_200, 10), (120_000, 12_000, 5), (12_000, 12_000, 10)]:
    q = torch.nn.functional.normalize(torch.randn(nq, 8, generator=g)).to(DEV)
    r = torch.nn.functional.normalize(torch.randn(nr, 8, generator=g)).to(DEV)
    t = timeit(lambda: ops.knn_radius(q, r, k, 1.5))
    print(f"| `hgnn_knn_radius` (brute force, D=8) | {nq:,} x {nr:,}, k={k} | {t:.3f} | {nq * nr / t / 1e6:.1f} G pairs/s |")
E = 600_000
gr = torch.stack([torch.randint(0, 120_000, (E,), generator=g), torch.randint(0, 12_000, (E,), generator=g)]).to(DEV)
t = timeit(lambda: ops.symmetrize(gr, 120_000))
print(f"| `hgnn_symmetrize` | {E:,} edges | {t:.3f} | {E / t / 1e3:.0f} M edges/s |")
t = timeit(lambda: ops.connected_components(gr, 120_000))
print(f"| `hgnn_connected_components` | {E:,} edges, 120,000 vertices | {t:.3f} | {E / t / 1e3:.0f} M edges/s |")

print("\n## Fused edge step, forward + backward (edge-steps/s; destination-sorted edges, N = E/10)\n")
print("| latent | E | graph | path | ms/step | M edge-steps/s |\n|---:|---:|---|---|---:|---:|")
# BASELINE config 2: latent 32-256 x 1e5-1e7 edges (the largest sizes are capped by what the layer-wise paths keep alive)
for L, E, pl in [(128, 100_000, False), (128, 1_000_000, False), (128, 4_000_000, False), (128, 10_000_000, False),
                 (128, 1_000_000, True),
                 (64, 100_000, False), (64, 1_000_000, False), (64, 10_000_000, False),
                 (32, 100_000, False), (32, 1_000_000, False), (32, 10_000_000, False),
                 (256, 100_000, False), (256, 1_000_000, False), (256, 4_000_000, False)]:
    hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
    torch.manual_seed(0)
    cell = InteractionGNNCell(hp); kaiming_init(cell); cell.to(DEV)
    n, e, gph = synth_edge_problem(E, L, seed=42, power_law=pl)
    order = torch.argsort(gph[1], stable=True); gph, e = gph[:, order].contiguous(), e[order].contiguous()
    n, e, gph = n.to(DEV).requires_grad_(True), e.to(DEV).requires_grad_(True), gph.to(DEV)
    N = n.shape[0]
    gp = GraphPlans(gph, N, N, dst_sorted=True); gp.by_src; gp.by_dst
    ce, ca = torch.randn_like(e), torch.randn(N, L, device=DEV)
    params = list(cell.edge_network.parameters())
    c0, r0 = ops.TC_CALLS["count"], ops.TC_ROW_CALLS["count"]
    def step():
        e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
        if agg is None:
            agg = ops.scatter_add(e2, gph[1], dim_size=N, plan=gp.by_dst)
        torch.autograd.grad([e2, agg], [n, e] + params, [ce, ca])
    t = timeit(step, n=5)
    used, used_rows = ops.TC_CALLS["count"] - c0, ops.TC_ROW_CALLS["count"] - r0
    path = ("fused tcgen05 edge kernels (fwd + bwd)" if used else
            ("layer-wise tcgen05 GEMMs + LayerNorm kernels (fwd + bwd)" if used_rows else "fp32 SIMT"))
    print(f"| {L} | {E:,} | {'power-law (hubs)' if pl else 'uniform'} | {path} | {t:.3f} | {E / t / 1e3:.1f} |")
    del cell, n, e, gph, gp

print("\n## Whole models, one synthetic event, forward + backward (default path)\n")
print("| model | event | ms | M edge-steps/s |\n|---|---|---:|---:|")
for name, npart, fpt in [("1 GeV", 1200, 4.0), ("full pile-up", 12000, 11.5)]:
    ev = synth_event(npart, 10, 0.0, fpt, seed=1000)
    x, gph = ev.x.to(DEV), ev.edge_index.to(DEV)
    Ed = 2 * gph.shape[1]
    torch.manual_seed(0)
    ec = model_selector("EC-IN"); kaiming_init(ec); ec.to(DEV)
    y = ev.y_pid.float().to(DEV)
    def fb():
        ec.zero_grad(set_to_none=True)
        torch.nn.functional.binary_cross_entropy(ec(x.clone(), gph), y).backward()
    t = timeit(fb, n=5, warm=4)
    print(f"| EC-IN (14 cells, latent 128) | {name}: N={x.shape[0]:,} E_d={Ed:,} | {t:.2f} | {Ed * 14 / t / 1e3:.1f} |")
    del ec
    bc = model_selector("4", dict(latent=128)); kaiming_init(bc); bc.to(DEV).train()
    clusters = (ev.pid - 1).to(DEV)
    def fb2():
        bc.zero_grad(set_to_none=True)
        bg, sc, emb = bc(x.clone(), gph, clusters=clusters)
        (sc.sum() + emb.sum()).backward()
    t = timeit(fb2, n=5, warm=4)
    print(f"| BC-HGNN-GMM (6+6 cells, latent 128, S={npart:,}) | {name}: N={x.shape[0]:,} E_d={Ed:,} | {t:.2f} | {Ed * 12 / t / 1e3:.1f} |")
    del bc
    torch.cuda.empty_cache()
