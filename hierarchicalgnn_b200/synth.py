"""Synthetic TrackML-shaped events and microbenchmark inputs.

All randomness comes from a CPU ``torch.Generator`` seeded explicitly, so the
oracle, the CUDA path and the CPU baseline see bit-identical inputs. Field
names follow the event layout documented in the reference dataset
(Modules/utils.py:38-51): ``x``, ``pid``, ``pt``, ``edge_index``, ``y_pid``.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch


def synth_event(n_particles=1200, hits_per_particle=10, noise_frac=0.0,
                fake_per_true=4.0, seed=1000):
    """One event: straight-ish tracks from the origin plus random fake edges.

    1 GeV shape: 1200 x 10 hits, fake_per_true=4  -> N=12000, E~54000.
    Full pile-up shape: 12000 x 10 hits, fake_per_true~13 -> N=120000, E~1.5M.
    """
    g = torch.Generator().manual_seed(int(seed))
    hpp = int(hits_per_particle)
    direction = torch.randn(n_particles, 3, generator=g)
    direction = direction / direction.norm(dim=1, keepdim=True)
    radii = torch.linspace(0.03, 1.0, hpp)
    x = direction[:, None, :] * radii[None, :, None]
    x = x + 0.002 * torch.randn(n_particles, hpp, 3, generator=g)
    x = x.reshape(-1, 3).float()
    pid = torch.arange(1, n_particles + 1).repeat_interleave(hpp)
    pt = 1.0 + torch.empty(n_particles).exponential_(1.0, generator=g)
    pt = pt.repeat_interleave(hpp)

    n_noise = int(round(noise_frac * x.shape[0]))
    if n_noise:
        xn = torch.randn(n_noise, 3, generator=g)
        xn = xn / xn.norm(dim=1, keepdim=True) * torch.rand(n_noise, 1, generator=g)
        x = torch.cat([x, xn.float()], 0)
        pid = torch.cat([pid, torch.zeros(n_noise, dtype=torch.long)])
        pt = torch.cat([pt, torch.zeros(n_noise)])

    n_hits = x.shape[0]
    base = torch.arange(n_particles * hpp).reshape(n_particles, hpp)
    true_edges = torch.stack([base[:, :-1].reshape(-1), base[:, 1:].reshape(-1)], 0)
    n_fake = int(round(fake_per_true * true_edges.shape[1]))
    fs = torch.randint(0, n_hits, (n_fake,), generator=g)
    fd = torch.randint(0, n_hits, (n_fake,), generator=g)
    keep = fs != fd
    fake_edges = torch.stack([fs[keep], fd[keep]], 0)
    edge_index = torch.cat([true_edges, fake_edges], 1)
    shuffle = torch.randperm(edge_index.shape[1], generator=g)
    edge_index = edge_index[:, shuffle].contiguous()
    y_pid = (pid[edge_index[0]] == pid[edge_index[1]]) & (pid[edge_index[0]] != 0)
    return SimpleNamespace(x=x.contiguous(), pid=pid, pt=pt.float(), edge_index=edge_index,
                           y_pid=y_pid, y=y_pid.clone(), n_particles=n_particles)


def synth_edge_problem(n_edges, latent, seed=42, nodes_per_edge=0.1, power_law=False):
    """Config-2 microbenchmark input: random node/edge latents on a symmetric
    directed graph (edge k and k+E/2 are mutual reverses), N = E * nodes_per_edge."""
    g = torch.Generator().manual_seed(int(seed))
    half = n_edges // 2
    n_nodes = max(2, int(round(n_edges * nodes_per_edge)))
    if power_law:
        u = torch.rand(half, generator=g)
        src = (n_nodes * u.pow(3.0)).long().clamp_(max=n_nodes - 1)
    else:
        src = torch.randint(0, n_nodes, (half,), generator=g)
    dst = torch.randint(0, n_nodes, (half,), generator=g)
    graph = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])], 0)
    nodes = torch.randn(n_nodes, latent, generator=g)
    edges = torch.randn(2 * half, latent, generator=g)
    return nodes, edges, graph


def direction_embeddings(event, dim=8, noise=0.02, seed=0):
    """Unit-norm ``dim``-d embeddings clustered by particle (used to exercise the
    supergraph builder with a realistic number of supernodes)."""
    g = torch.Generator().manual_seed(int(seed))
    centres = torch.randn(event.n_particles + 1, dim, generator=g)
    emb = centres[event.pid] + noise * torch.randn(event.x.shape[0], dim, generator=g)
    return torch.nn.functional.normalize(emb).float()


def collate_events(events):
    """Several events as one disjoint graph, laid out like a torch_geometric ``Batch``: rows event by event, ``edge_index``
    offset by the hits before it, ``batch`` = event id per hit, ``ptr`` = hit offsets, ``num_graphs``. ``pid`` stays the
    per-event particle id (as torch_geometric leaves it); ``clusters`` (supernode = particle, pid - 1, offset by the
    particles of the earlier events; -1 for noise) is what the benchmarks inject in place of the learned clustering."""
    xs, gs, pids, pts, ys, bs, cl = [], [], [], [], [], [], []
    n_hits = n_part = 0
    ptr = [0]
    for b, ev in enumerate(events):
        xs.append(ev.x)
        gs.append(ev.edge_index + n_hits)
        pids.append(ev.pid)
        pts.append(ev.pt)
        ys.append(ev.y_pid)
        bs.append(torch.full((ev.x.shape[0],), b, dtype=torch.long))
        c = ev.pid - 1
        cl.append(torch.where(c >= 0, c + n_part, c))
        n_hits += ev.x.shape[0]
        n_part += ev.n_particles
        ptr.append(n_hits)
    y = torch.cat(ys)
    return SimpleNamespace(x=torch.cat(xs).contiguous(), edge_index=torch.cat(gs, 1).contiguous(), pid=torch.cat(pids),
                           pt=torch.cat(pts), y_pid=y, y=y.clone(), batch=torch.cat(bs), ptr=torch.tensor(ptr),
                           num_graphs=len(events), clusters=torch.cat(cl), n_particles=n_part)
