"""torch.profiler view of one BC-HGNN-GMM (latent 128) fwd+bwd on a synthetic 1 GeV event: which kernels / host ops
are left around the hgnn_b200 launches (torch glue, host syncs)."""
import sys, torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector

which = sys.argv[1] if len(sys.argv) > 1 else "bc"
ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
x, g = ev.x.cuda(), ev.edge_index.cuda()
torch.manual_seed(0)
if which == "bc":
    m = model_selector("4", dict(latent=128)); kaiming_init(m); m.cuda().train()
    clusters = (ev.pid - 1).cuda()
    def fb():
        m.zero_grad(set_to_none=True)
        bg, sc, emb = m(x.clone(), g, clusters=clusters)
        (sc.sum() + emb.sum()).backward()
else:
    m = model_selector("EC-IN"); kaiming_init(m); m.cuda()
    y = ev.y_pid.float().cuda()
    def fb():
        m.zero_grad(set_to_none=True)
        torch.nn.functional.binary_cross_entropy(m(x.clone(), g), y).backward()
for _ in range(3): fb()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
