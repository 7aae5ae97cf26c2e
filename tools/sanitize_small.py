"""small invocations of every kernel for compute-sanitizer (memcheck): ragged sizes, both modes, both dtypes"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppo_rl_satellite_b200 import engine as eng
import bench
rng = np.random.default_rng(0)
for n in (1, 33, 130):
    for mode in ("cw", "rk4"):
        env = eng.EnvBatch(n, mode=mode, substeps=3, d_capture=181200.0, max_episode_steps=3)
        st, rs = eng.RunningStats(18), eng.RunningStats(1)
        obs32 = torch.empty((n, 18), dtype=torch.float32, device="cuda"); obs64 = torch.empty((n, 18), dtype=torch.float64, device="cuda")
        term = torch.empty((n, 18), dtype=torch.float64, device="cuda"); std = torch.zeros(1, dtype=torch.float64, device="cuda")
        for t in range(5):
            dt = torch.float32 if t % 2 else torch.float64
            pa = torch.as_tensor(rng.uniform(-2, 2, (n, 3)), dtype=dt, device="cuda"); ea = torch.as_tensor(rng.uniform(-2, 2, (n, 3)), dtype=dt, device="cuda")
            env.step(pa, ea, obs_f32=obs32, obs_f64=obs64, term_obs_f64=term, obs_stats=st, ret_stats=rs, ret_std_out=std)
        env.step_host(rng.uniform(-2, 2, (n, 3)).astype(np.float32), rng.uniform(-2, 2, (n, 3)).astype(np.float32))
        env.reset(mask=env.done); env.observe(torch.float32); env.observe(torch.float64)
        actor = eng.GaussianActorKernel().load_state_dict(bench.orthogonal_actor_state(torch, 0))
        actor.sample(env=env, obs_stats=st, seed=1, step=2, obs_out=obs32, mean_out=torch.empty((n, 3), device="cuda"), eps_out=torch.empty((n, 3), device="cuda"))
        actor.sample(obs=obs32, eps_in=torch.zeros((n, 3), device="cuda"))
    x, _ = eng.alloc_soa(6, n, torch.float64, "cuda")
    x[0] = 7000.0; x[4] = 7.5
    eng.rk4_propagate(x, 1.0, 5); eng.rk4_propagate(x, 1.0, 5, j2=0.0)
    eng.Rk4HostPropagator(n)(x.cpu().numpy(), 1.0, 3)
    T = 7
    r = torch.randn((T, n), device="cuda"); v = torch.randn((T + 1, n), device="cuda"); d = (torch.rand((T, n), device="cuda") < 0.2).to(torch.uint8)
    adv, vt = eng.gae_time_major(r, v, d, r_scale=torch.rand(T, device="cuda"))
    eng.adv_normalize_(adv, group=False) if adv.numel() > 1 else None
    f = torch.randn(n * 50, device="cuda")
    eng.gae_flat(f, f, f, (f > 1).float(), (f > 1).float())
    rv = torch.randn((n, 12), dtype=torch.float64, device="cuda") * 10 + torch.tensor([27298000.0, 32306000.0, 0, -2350.0, 1970.0, 0.1, 27116000.0, 32306000.0, 10, -2350.0, 1970.0, 0.2], dtype=torch.float64, device="cuda")
    eng.danger_zone_count(rv, torch.full((n,), 300.0, dtype=torch.float64, device="cuda"), debug=True)
x, _ = eng.alloc_soa(6, 151552 + 3, torch.float64, "cuda"); x[0] = 7000.0; x[4] = 7.5
eng.rk4_propagate(x, 1.0, 2)
critic = eng.GaussianActorKernel(critic=True)
sd = bench.orthogonal_actor_state(torch, 2)
critic.load_state_dict({"fc1.weight": sd["fc1.weight"], "fc1.bias": sd["fc1.bias"], "fc2.weight": sd["fc2.weight"], "fc2.bias": sd["fc2.bias"],
                        "fc3.weight": torch.randn(1, 256), "fc3.bias": torch.zeros(1)})
critic.value(torch.randn(70, 18, device="cuda"))
sys.path.insert(0, os.path.join(ROOT, "ppo-rl-satellite_b200", "dropin"))
import satellite_function as sf, orbit_rk4
sf.Numerical_calculation_method(np.array([1e5, 2e4, 3e3]), np.array([1.0, 2, 3]), np.array([1e3, 2e3, 3e2]), np.array([0.1, 0.2, 0.3])).numerical_calculation(100)
sf.Clohessy_Wiltshire(np.ones(3), np.ones(3), np.ones(3), np.ones(3)).State_transition_matrix(100)
orbit_rk4.StateEq(0, np.array([7000.0, 0, 0, 0, 7.5, 0]))
sf.state_information_batch(np.array([[42164000.0, 0.001, 0.1, 1, 2, 3]]))
torch.cuda.synchronize()
print("sanitize_small done")
