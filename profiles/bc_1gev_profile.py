"""Where one BC-HGNN-GMM forward + backward on a 1 GeV event (BASELINE config 3: the e2e line of bench.py) spends its time:
wall per step, GPU busy time per step (sum of kernel durations), launches per step, top kernels and top host-side ops.
Usage: python profiles/bc_1gev_profile.py [events per step: > 1 = collated into one disjoint graph]"""
import sys, time, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200.synth import collate_events, synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ev = collate_events([synth_event(1200, 10, 0.0, 4.0, seed=1000 + i) for i in range(B)])
x, g = ev.x.cuda(), ev.edge_index.cuda()
bt = ev.batch.cuda() if B > 1 else None
torch.manual_seed(0)
import os
LAT = int(os.environ.get("HGNN_LATENT", "128"))  # 256 = the reference's BC default (layer-wise tensor-core path)
bc = model_selector("BC-HGNN-GMM", dict(latent=LAT)); kaiming_init(bc); bc.cuda().train()
clusters = ev.clusters.cuda()
def fb(split=False):
    bc.zero_grad(set_to_none=True)
    bg, sc, emb = bc(x.clone(), g, clusters=clusters, batch=bt, n_events=B if B > 1 else None)
    if split: torch.cuda.synchronize(); t = time.perf_counter()
    (sc.sum() + emb.sum()).backward()
    return t if split else None
for _ in range(5): fb()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): fb()
torch.cuda.synchronize(); print(f"{B} event(s) per step, latent {LAT}"); print("BC 1 GeV fwd+bwd %.2f ms/step (back to back)" % ((time.perf_counter() - t0) / 10 * 1e3))
tf = tb = 0.0
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); t1 = fb(True); torch.cuda.synchronize(); t2 = time.perf_counter()
    tf += t1 - t0; tb += t2 - t1
print("forward %.2f ms, backward %.2f ms" % (tf / 5 * 1e3, tb / 5 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    for _ in range(3): fb()
    torch.cuda.synchronize()
ka = p.key_averages()
kern = [e for e in ka if e.device_type == torch.autograd.DeviceType.CUDA]
print("GPU busy %.2f ms/step in %d launches/step" % (sum(e.device_time_total for e in kern) / 3e3, sum(e.count for e in kern) // 3))
print(ka.table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
print(ka.table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
