"""Event-timed forward / backward of the fan-in <= 8 first layers (hgnn_narrow_in_*): fan-in 3 on hit rows, fan-in 6 gathered
from the two end points of an edge, fan-out 256 / 512, 1 M rows."""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.utils import make_mlp
torch.manual_seed(0)
rows = 1_000_000
for fan_out in (256, 512):
    for gathered in (False, True):
        mlp = make_mlp(6 if gathered else 3, fan_out, fan_out, 1, layer_norm=True, output_activation="GELU", hidden_activation="GELU").cuda()
        x = torch.randn(100_000 if gathered else rows, 3, device="cuda", requires_grad=True)
        if gathered:
            idx = [torch.randint(0, x.shape[0], (rows,), device="cuda") for _ in range(2)]
            plans = [ops.plan_for(i, x.shape[0]) for i in idx]
            f = lambda: mlp.fused([x, x], plans)
        else:
            f = lambda: mlp.fused([x], [None])
        cot = torch.randn(rows, fan_out, device="cuda")
        params = list(mlp.parameters())
        def fb():
            y = f()
            torch.autograd.grad(y, [x] + params, cot)
        for fn, name in ((lambda: f().detach(), "forward"), (fb, "forward + backward")):
            for _ in range(3): fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(10): fn()
            b.record(); torch.cuda.synchronize()
            print(f"fan-in {6 if gathered else 3}{' (gathered)' if gathered else ''} -> {fan_out}, {rows} rows: {name} {a.elapsed_time(b) / 10:.3f} ms")
