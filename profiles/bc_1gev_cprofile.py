"""Host-side (Python) profile of the BC-HGNN-GMM forward + backward on a 1 GeV event: the step is host-bound at this event
size (GPU busy 12.7 ms of 18.1 ms), so this is the list that matters for bench.py's e2e line. cProfile over 10 steps."""
import cProfile, pstats, sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
x, g = ev.x.cuda(), ev.edge_index.cuda()
torch.manual_seed(0)
bc = model_selector("BC-HGNN-GMM", dict(latent=128)); kaiming_init(bc); bc.cuda().train()
clusters = (ev.pid - 1).cuda()
def fb():
    bc.zero_grad(set_to_none=True)
    bg, sc, emb = bc(x.clone(), g, clusters=clusters)
    (sc.sum() + emb.sum()).backward()
for _ in range(5): fb()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10): fb()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(60)
