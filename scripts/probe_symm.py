"""Probe (2+ GPUs): is NVSwitch multicast reachable through torch symmetric memory on this box?"""
import os
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, group=dist.group.WORLD)
    if rank == 0:
        print("symm ok: multicast_ptr=%#x buffer_ptrs=%s signal_pads=%d" % (h.multicast_ptr, [hex(p) for p in h.buffer_ptrs], len(h.signal_pad_ptrs)), flush=True)
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print("symm failed:", repr(e)[:400], flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1), flush=True)
x = torch.ones(10_300_000 // 4 * 1, device=dev)
for n in (2_575_000, 10_532_000):
    x = torch.ones(n, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"nccl all_reduce {n * 4 / 1e6:.1f} MB: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
dist.destroy_process_group()
