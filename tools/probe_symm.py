"""probe: torch symmetric memory on this box (peer pointers, device barrier cost) vs a small NCCL all-reduce"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
N = 71432
t = symm_mem.empty(N, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
if rank == 0:
    print("ptrs", [hex(p) for p in hdl.buffer_ptrs], "mc_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "signal pad", hdl.signal_pad_size, flush=True)
t.fill_(float(rank + 1))
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (N,), torch.float32)
print(rank, "peer value", float(peer[0]), float(peer[-1]), flush=True)
hdl.barrier()
x = torch.ones(N, device=dev)
for name, fn in (("symm barrier", lambda: hdl.barrier()), ("nccl all_reduce 286KB", lambda: dist.all_reduce(x))):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us", flush=True)
dist.destroy_process_group()
