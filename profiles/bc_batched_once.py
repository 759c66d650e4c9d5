"""Two forward + backward steps of BC-HGNN-GMM (latent 128) on a collated batch of 1 GeV events — the workload of bench.py's
`e2e` — for an ncu launch list:  ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file L.csv \\
    python profiles/bc_batched_once.py [events per step]"""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200.synth import collate_events, synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ev = collate_events([synth_event(1200, 10, 0.0, 4.0, seed=1000 + i) for i in range(B)])
x, g, bt, cl = ev.x.cuda(), ev.edge_index.cuda(), ev.batch.cuda(), ev.clusters.cuda()
torch.manual_seed(0)
bc = model_selector("BC-HGNN-GMM", dict(latent=128)); kaiming_init(bc); bc.cuda().train()
for _ in range(2):
    bc.zero_grad(set_to_none=True)
    bg, sc, emb = bc(x.clone(), g, clusters=cl, batch=bt, n_events=B)
    (sc.mean() + emb.square().mean()).backward()
torch.cuda.synchronize()
print("ok")
