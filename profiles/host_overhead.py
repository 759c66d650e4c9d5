"""cProfile view of the host side of one BC-HGNN-GMM / EC-IN forward+backward (1 GeV event): where the Python time between
kernel launches goes. Usage: python profiles/host_overhead.py [bc|ec]"""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector

which = sys.argv[1] if len(sys.argv) > 1 else "bc"
fresh = len(sys.argv) > 2 and sys.argv[2] == "fresh"   # a new edge-list tensor every step (as a data loader delivers): no cached plans
ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
x, g = ev.x.cuda(), ev.edge_index.cuda()
torch.manual_seed(0)
if which == "bc":
    m = model_selector("4", dict(latent=128)); kaiming_init(m); m.cuda().train()
    clusters = (ev.pid - 1).cuda()
    def step():
        m.zero_grad(set_to_none=True)
        bg, sc, emb = m(x.clone(), g.clone() if fresh else g, clusters=clusters)
        (sc.sum() + emb.sum()).backward()
else:
    m = model_selector("EC-IN"); kaiming_init(m); m.cuda()
    y = ev.y_pid.float().cuda()
    def step():
        m.zero_grad(set_to_none=True)
        torch.nn.functional.binary_cross_entropy(m(x.clone(), g.clone() if fresh else g), y).backward()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print(f"{which}: wall {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/step, launches/step {ops.LAUNCHES['count'] // 8}")
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
