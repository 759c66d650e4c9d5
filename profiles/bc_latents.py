import sys, time, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
x, g = ev.x.cuda(), ev.edge_index.cuda()
Ed = 2 * g.shape[1]
clusters = (ev.pid - 1).cuda()
for lat in (256, 64):
    torch.manual_seed(0)
    bc = model_selector("4", dict(latent=lat)); kaiming_init(bc); bc.cuda().train()
    for mode in ("fp32", "auto"):
        ops.set_precision(mode)
        def fb():
            bc.zero_grad(set_to_none=True)
            bg, sc, emb = bc(x.clone(), g, clusters=clusters)
            (sc.sum() + emb.sum()).backward()
        for _ in range(2): fb()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): fb()
        torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 3 * 1e3
        print(f"BC-HGNN latent {lat} [{mode}]: fwd+bwd {t:.2f} ms ({Ed*12/t/1e3:.1f} M edge-steps/s)")
    ops.PROFILE = {}
    fb(); torch.cuda.synchronize()
    prof = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in ops.PROFILE.items()}
    ops.PROFILE = None
    for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]:
        print(f"   {k:24s} calls {n:4d}  total {t:8.2f} ms")
