"""2+ GPUs (torchrun): the fused PPO step with the gradient exchange inside the Adam kernel (NVLink peer memory) against the
same step with NCCL all-reduce: parameters after 2 x K epochs, replica consistency, and the time per optimiser step."""
import os, sys, time, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ppo_rl_satellite_b200.dropin import ppo_continuous as P
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B, mb = 1 << 19, 65536
a = bench._PpoArgs(policy_dist="Gaussian", max_action=1.6, batch_size=B, mini_batch_size=mb, max_train_steps=int(3e6),
                   lr_a=2e-4, lr_c=2e-4, gamma=0.99, lamda=0.95, epsilon=0.1, K_epochs=2, entropy_coef=0.01, set_adam_eps=True,
                   use_grad_clip=True, use_lr_decay=True, use_adv_norm=True, state_dim=18, action_dim=3, hidden_width=256,
                   use_tanh=True, use_orthogonal_init=True, chkpt_dir="/tmp")
g = torch.Generator(device=dev).manual_seed(100 + rank)                 # different data per rank
s = torch.randn(B, 18, device=dev, generator=g); act = torch.randn(B, 3, device=dev, generator=g).clamp(-1.6, 1.6)
lp = torch.randn(B, 3, device=dev, generator=g) * 0.1 - 1.0
adv = torch.randn(B, 1, device=dev, generator=g); vt = torch.randn(B, 1, device=dev, generator=g)
out = {}
for mode in ("nccl", "peer"):
    os.environ["SAT_PEER_ALLREDUCE"] = "1" if mode == "peer" else "0"
    torch.manual_seed(0)
    agent = P.PPO_continuous(a, "pursuer", device=dev)
    torch.manual_seed(7)
    agent.optimize(s, act, lp, adv, vt, mini_batch_size=mb, fused=True)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    agent.optimize(s, act, lp, adv, vt, mini_batch_size=mb, fused=True)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    assert any(agent._fused["peers"].values()) == (mode == "peer")
    flat = torch.cat([p.detach().reshape(-1) for p in list(agent.actor.parameters()) + list(agent.critic.parameters())])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    out[mode] = flat
    if rank == 0:
        print(f"{mode}: {1e3 * dt / (2 * B // mb):.3f} ms per optimiser step, replicas bit-identical: {same}", flush=True)
if rank == 0:
    d = (out["nccl"] - out["peer"]).abs()
    print(f"peer vs nccl parameters: max |diff| {d.max().item():.3e}, mean {d.mean().item():.3e}", flush=True)
dist.destroy_process_group()
