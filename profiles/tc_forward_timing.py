import sys, torch, time
sys.path.insert(0,'.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
for L,E in [(128,1000000),(64,1000000),(128,120000)]:
    hp=dict(latent=L,hidden=2*L,nb_edge_layer=2,nb_node_layer=3,layernorm=True,hidden_activation="GELU")
    torch.manual_seed(0); cell=InteractionGNNCell(hp); kaiming_init(cell); cell.cuda()
    n,e,g=synth_edge_problem(E,L); n,e,g=n.cuda(),e.cuda(),g.cuda()
    gp=GraphPlans(g,n.shape[0],n.shape[0]); gp.by_src; gp.by_dst
    for mode in ("bf16","fp32"):
        ops.set_precision(mode)
        with torch.no_grad():
            for _ in range(3): cell.edge_update(n,e,gp)
            torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): cell.edge_update(n,e,gp)
            b.record(); torch.cuda.synchronize()
        ms=a.elapsed_time(b)/10
        print(f"L={L} E={E} {mode}: {ms:.3f} ms/fwd  {E/ms/1e6:.3f} G edges/s  hbm_frac={(8*L+8+0.8*L)*E/ms/1e6/6554:.3f} tensor_frac={16*L*L*E/ms/1e9/1350:.3f}")
