"""Wall-clock breakdown of one BC-HGNN-GMM training step (1 GeV events, one GPU): forward, embedding loss, assignment loss
(host-side matching), backward, clip, optimizer. Each part is bracketed by synchronize (so parts do not overlap).
Usage: python profiles/train_step_breakdown.py [events per step, default 1: collated into one disjoint graph when > 1]"""
import sys, time, torch
from types import SimpleNamespace
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.synth import collate_events, synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
from hierarchicalgnn_b200.parallel import clip_grad_norm_
dev = 'cuda'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ev = collate_events([synth_event(1200, 10, 0.0, 4.0, seed=1000 + i) for i in range(B)])
b = SimpleNamespace(x=ev.x.to(dev), edge_index=ev.edge_index.to(dev), pid=ev.pid.to(dev), pt=ev.pt.to(dev))
if B > 1:
    b.batch, b.num_graphs = ev.batch.to(dev), B
torch.manual_seed(0)
m = model_selector("BC-HGNN-GMM", dict(latent=128, loss_schedule=0.5)); kaiming_init(m); m.to(dev).train()
clusters = ev.clusters.to(dev)
m.hgnn_block.clustering = lambda x, emb, graph: clusters
opt = m.configure_optimizers()[0][0]
params = list(m.parameters())
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return time.perf_counter()
def step(measure):
    t = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    bg, sc, emb = m(b.x.clone(), b.edge_index, batch=getattr(b, "batch", None), n_events=B if B > 1 else None)
    if measure: t = tick("forward", t)
    el = m.embedding_loss(b, emb)
    if measure: t = tick("embedding loss", t)
    al = m.assignment_loss(b, bg, sc)
    if measure: t = tick("assignment loss (host-side matching)      ", t)
    (0.5 * el + 0.5 * al).backward()
    if measure: t = tick("backward", t)
    clip_grad_norm_(params, 0.5)
    if measure: t = tick("clip", t)
    opt.step()
    if measure: t = tick("optimizer", t)
for _ in range(3): step(False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step(False)
torch.cuda.synchronize(); print(f"{B} event(s) per step, unbracketed: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/step")
for _ in range(5): step(True)
for k, v in T.items(): print(f"  {k:45s} {v / 5 * 1e3:7.2f} ms")
