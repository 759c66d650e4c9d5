"""Run-to-run bit-identity of the BC config-3 forward + backward on the default (tensor-core) path, with the caching
allocator's free memory poisoned (NaN bit patterns) between runs: a kernel that reads a workspace it never wrote, or that
races, shows up as a changed output / gradient or a NaN. (A home-made initcheck: compute-sanitizer is not available on the pool.)"""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200.synth import synth_event
from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
dev = 'cuda'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
torch.manual_seed(0)
model = model_selector("BC-HGNN-GMM", dict(latent=128)); kaiming_init(model); model.to(dev).train()
ev = synth_event(1200, 10, 0.0, 4.0, seed=1000)
clusters = (ev.pid - 1).to(dev)
x, ei = ev.x.to(dev), ev.edge_index.to(dev)
g = torch.Generator().manual_seed(5)
ws = we = None
def poison(pattern):
    # fill what the allocator holds free with a bit pattern, then hand it back to the cache (not to the driver)
    free = torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
    blocks = []
    for sz in (free // 2, free // 4, free // 8, 64 << 20, 16 << 20, 4 << 20, 1 << 20, 1 << 18):
        for _ in range(4):
            try:
                b = torch.empty(max(sz, 1024) // 4, dtype=torch.int32, device=dev)
            except RuntimeError:
                break
            b.fill_(pattern); blocks.append(b)
    del blocks
    torch.cuda.synchronize()
def run():
    global ws, we
    for p in model.parameters(): p.grad = None
    bg, scores, emb = model(x.clone(), ei, clusters=clusters)
    if ws is None:
        ws, we = torch.randn(scores.shape, generator=g).to(dev), torch.randn(emb.shape, generator=g).to(dev)
    ((scores * ws).sum() + (emb * we).sum()).backward()
    out = {"bg": bg.clone(), "scores": scores.detach().clone(), "emb": emb.detach().clone()}
    for k, p in model.named_parameters():
        if p.grad is not None: out["d." + k] = p.grad.clone()
    return out
ref = run()
bad = 0
for it in range(N):
    poison([0x7fc00000, 0x7f800001, -1, 0x7f7fffff, 0x12345678][it % 5] if it % 5 != 2 else -1)
    cur = run()
    for k in ref:
        if cur[k].shape != ref[k].shape or not torch.equal(cur[k], ref[k]):
            bad += 1
            d = (cur[k].double() - ref[k].double()).abs().max().item() if cur[k].shape == ref[k].shape else float('nan')
            print(f"run {it}: {k} differs (max abs {d:.3e}, nan {bool(torch.isnan(cur[k].float()).any())})")
print(f"{N} poisoned re-runs, {len(ref)} tensors each: {bad} mismatches")
