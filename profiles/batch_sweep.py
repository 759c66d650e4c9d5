"""BC-HGNN-GMM (latent 128) forward + backward on collated batches of 1 GeV events: ms per step and per event against the
number of events per step (where the step turns from host-bound to GPU-bound)."""
import sys, time, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200.synth import collate_events, synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
torch.manual_seed(0)
bc = model_selector("BC-HGNN-GMM", dict(latent=128)); kaiming_init(bc); bc.cuda().train()
events = [synth_event(1200, 10, 0.0, 4.0, seed=1000 + i) for i in range(32)]
print("| events per step | ms per step | ms per event | M edge-steps/s |\n|---:|---:|---:|---:|")
for B in (1, 2, 4, 8, 16, 32):
    ev = collate_events(events[:B])
    x, g, bt, cl = ev.x.cuda(), ev.edge_index.cuda(), ev.batch.cuda(), ev.clusters.cuda()
    def step():
        for p in bc.parameters(): p.grad = None
        bg, sc, emb = bc(x.clone(), g, clusters=cl, batch=bt if B > 1 else None, n_events=B if B > 1 else None)
        (sc.mean() + emb.square().mean()).backward()
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 8 if B <= 8 else 4
    for _ in range(n): step()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / n * 1e3
    es = 2 * ev.edge_index.shape[1] * 12
    print(f"| {B} | {ms:.1f} | {ms / B:.2f} | {es / ms / 1e3:.0f} |")
    del x, g, bt, cl
    torch.cuda.empty_cache()
