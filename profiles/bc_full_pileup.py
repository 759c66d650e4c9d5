import sys, time, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
ev = synth_event(12000, 10, 0.0, 11.5, seed=1000)
x, g = ev.x.cuda(), ev.edge_index.cuda()
bc = model_selector("4", dict(latent=128)); kaiming_init(bc); bc.cuda().train()
clusters = (ev.pid - 1).cuda()
def fb():
    bc.zero_grad(set_to_none=True)
    bg, sc, emb = bc(x.clone(), g, clusters=clusters)
    (sc.sum() + emb.sum()).backward()
for _ in range(3): fb()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): fb()
torch.cuda.synchronize(); print("BC full pile-up fwd+bwd %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
ops.PROFILE = {}
fb(); torch.cuda.synchronize()
prof = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in ops.PROFILE.items()}
ops.PROFILE = None
print("sum %.2f" % sum(t for _, t in prof.values()))
for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"   {k:24s} calls {n:4d}  total {t:8.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    fb(); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
