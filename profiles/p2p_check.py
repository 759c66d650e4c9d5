"""Multi-GPU check + timing of the peer-memory row collectives against NCCL (torchrun, one process per GPU).
torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/p2p_check.py [block] [width]"""
import os, sys, time, torch
import torch.distributed as dist
sys.path.insert(0, '.')
from hierarchicalgnn_b200.parallel import SymmetricRows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
block = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(rank)
x = torch.randn(block, width, device=dev)
full = torch.randn(world * block, width, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


want_ag = torch.empty(world * block, width, device=dev)
dist.all_gather_into_tensor(want_ag, x)
want_rs = torch.empty(block, width, device=dev)
dist.reduce_scatter_tensor(want_rs, full.clone())
for mc in (True, False):
    try:
        sr = SymmetricRows(block, width, dev, use_multicast=mc)
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            print("SymmetricRows failed:", repr(ex))
        raise
    got_ag = sr.all_gather(x)
    got_rs = sr.reduce_scatter(full)
    torch.cuda.synchronize()
    ok_ag = torch.equal(got_ag, want_ag)
    err_rs = float((got_rs - want_rs).abs().max())
    flat = torch.randn(350_001, device=dev)
    want_ar = flat.clone(); dist.all_reduce(want_ar)
    got_ar = sr.all_reduce_(flat.clone())
    err_ar = float((got_ar - want_ar).abs().max())
    t_ar = timeit(lambda: sr.all_reduce_(flat))
    t_ag, t_rs = timeit(lambda: sr.all_gather(x)), timeit(lambda: sr.reduce_scatter(full))
    if rank == 0:
        print(f"multicast requested {mc}, used {bool(sr.mc_base)}: all-gather equal {ok_ag}, reduce-scatter max |diff| vs NCCL {err_rs:.2e}; "
              f"all-reduce (1.4 MB) max |diff| {err_ar:.2e}, {t_ar * 1e3:.0f} us; all-gather {t_ag * 1e3:.0f} us, reduce-scatter {t_rs * 1e3:.0f} us  (table {world * block * width * 4 / 1e6:.1f} MB)")
    del sr
out = torch.empty_like(want_ag)
t_ag = timeit(lambda: dist.all_gather_into_tensor(out, x))
o2 = torch.empty_like(want_rs)
t_rs = timeit(lambda: dist.reduce_scatter_tensor(o2, full))
if rank == 0:
    print(f"NCCL: all-gather {t_ag * 1e3:.0f} us, reduce-scatter {t_rs * 1e3:.0f} us")
dist.destroy_process_group()
