"""What can this box's host side sustain? Every rank moves exactly the bytes EnvBatch.step_host moves per step at 65 536 envs
(H2D 1.57 MB of actions, D2H 5.31 MB of observation / reward / done) between PINNED host memory and its GPU with plain
cudaMemcpyAsync on two streams (copy engines only, no kernels), all ranks concurrently. Launch with torchrun at 1 / 2 / 4 / 8
ranks; rank 0 prints GB/s per GPU and the step rate this ceiling would allow. Used to tell whether the e2e scaling of
bench.py (0.54 at 8 GPUs in round 1) is the hardware or the code.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py"""
import os, sys, time
import torch, torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
if os.environ.get("SAT_NUMA_BIND", "1") != "0" and world > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    bench.bind_to_gpu_numa_node(lr)
n = 65536
h2d_bytes, d2h_bytes = n * 24, n * 81
h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev); d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def loop(iters, mode):
    for _ in range(iters):
        if mode in ("both", "h2d"):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("both", "d2h"):
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
        if mode == "step":                      # serialised like one env step: actions in, then results out, then sync
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            s_out.wait_stream(s_in)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
            s_out.synchronize()
    torch.cuda.synchronize()


res = {}
for mode in ("d2h", "h2d", "both", "step"):
    loop(20, mode)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    iters = 200
    loop(iters, mode)
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res[mode] = float(dt.item()) / iters
if rank == 0:
    print(f"ranks {world}: per GPU  D2H alone {d2h_bytes / res['d2h'] / 1e9:5.1f} GB/s   H2D alone {h2d_bytes / res['h2d'] / 1e9:5.1f} GB/s   "
          f"both directions concurrently {d2h_bytes / res['both'] / 1e9:5.1f} + {h2d_bytes / res['both'] / 1e9:4.1f} GB/s   "
          f"serialised step (H2D then D2H then sync) {res['step'] * 1e6:6.1f} us -> copy-only ceiling {n * world / res['step']:.3e} env-steps/s "
          f"({n / res['step']:.3e} per GPU)", flush=True)
if world > 1:
    dist.destroy_process_group()
