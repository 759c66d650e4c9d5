"""Model-level timing on one synthetic 1 GeV event (N=12000, E~54k -> 108k directed): EC-IN (config 1) forward and
fwd+bwd, BC-HGNN-GMM latent 128 (config 3) forward+backward with injected clusters; fp32 SIMT path vs default path."""
import sys, time, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
x, g = ev.x.cuda(), ev.edge_index.cuda()
Ed = 2 * g.shape[1]
torch.manual_seed(0)
ec = model_selector("EC-IN"); kaiming_init(ec); ec.cuda()
y = ev.y_pid.float().cuda()
for mode in ("fp32", "auto"):
    ops.set_precision(mode)
    with torch.no_grad():
        t_f = timeit(lambda: ec(x, g))
    def fb():
        ec.zero_grad(set_to_none=True)
        s = ec(x.clone(), g)
        torch.nn.functional.binary_cross_entropy(s, y).backward()
    t_fb = timeit(fb)
    print(f"EC-IN  [{mode:4s}] N={x.shape[0]} E_d={Ed}: fwd {t_f:8.2f} ms ({Ed*14/t_f/1e3:7.1f} M edge-steps/s)   fwd+bwd {t_fb:8.2f} ms ({Ed*14/t_fb/1e3:7.1f} M edge-steps/s)")
bc = model_selector("4", dict(latent=128)); kaiming_init(bc); bc.cuda().train()
clusters = (ev.pid - 1).cuda()
for mode in ("fp32", "auto"):
    ops.set_precision(mode)
    def fb():
        bc.zero_grad(set_to_none=True)
        bg, sc, emb = bc(x.clone(), g, clusters=clusters)
        (sc.sum() + emb.sum()).backward()
    t = timeit(fb, 3)
    print(f"BC-HGNN[{mode:4s}] latent 128, 6+6 cells, S={int(clusters.max())+1}: fwd+bwd {t:8.2f} ms ({Ed*12/t/1e3:7.1f} M edge-steps/s)")

# per-operator GPU time of one EC-IN fwd+bwd on the default path (CUDA events around each C-ABI call)
import statistics, time
ops.set_precision("auto")
ops.PROFILE = {}
torch.cuda.synchronize(); t0 = time.perf_counter()
ec.zero_grad(set_to_none=True)
s_ = ec(x.clone(), g)
torch.nn.functional.binary_cross_entropy(s_, y).backward()
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
prof = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in ops.PROFILE.items()}
ops.PROFILE = None
print(f"EC-IN fwd+bwd wall {wall:.1f} ms; per-op GPU ms:")
for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:24s} calls {n:4d}  total {t:8.2f} ms")

ops.PROFILE = {}
bc.zero_grad(set_to_none=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
bg, sc, emb = bc(x.clone(), g, clusters=clusters)
(sc.sum() + emb.sum()).backward()
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
prof = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in ops.PROFILE.items()}
ops.PROFILE = None
print(f"BC-HGNN fwd+bwd wall {wall:.1f} ms; per-op GPU ms (sum {sum(t for _, t in prof.values()):.2f}):")
for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:24s} calls {n:4d}  total {t:8.2f} ms")
