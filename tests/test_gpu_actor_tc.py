"""K3 on the tensor cores (csrc/actor_tc.cu): the exact-bf16x3 tcgen05 path of the actor's dense layers against an fp64 ground
truth, next to the fp32 FFMA2 path (csrc/actor.cu). The reference evaluates Actor_Gaussian in fp32
(ppo_continuous.py:83-95); the tensor-core path is the default because it is as close to the exact result as fp32 FMA is."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    from ppo_rl_satellite_b200 import engine
    return engine


def _weights(seed, gain3=1.0):
    import bench
    W = {k: v.numpy().copy() for k, v in bench.orthogonal_actor_state(torch, seed).items()}
    rng = np.random.default_rng(seed)
    W["fc1.bias"] = rng.normal(0, 0.1, 256).astype(np.float32)
    W["fc2.bias"] = rng.normal(0, 0.1, 256).astype(np.float32)
    W["mean_layer.weight"] = (W["mean_layer.weight"] * gain3).astype(np.float32)
    W["mean_layer.bias"] = rng.normal(0, 0.1, 3).astype(np.float32)
    W["log_std"] = np.array([[-0.3, 0.1, 0.2]], dtype=np.float32)
    return W


def _fp64_mean(W, x, use_tanh=True, max_action=1.6):
    act = np.tanh if use_tanh else (lambda v: np.maximum(v, 0.0))
    f = lambda k: W[k].astype(np.float64)
    h1 = act(x.astype(np.float64) @ f("fc1.weight").T + f("fc1.bias"))
    h2 = act(h1 @ f("fc2.weight").T + f("fc2.bias"))
    return max_action * np.tanh(h2 @ f("mean_layer.weight").T + f("mean_layer.bias"))


@pytest.mark.timeout(180)
@pytest.mark.parametrize("use_tanh", [True, False])
def test_tc_hidden_layer_is_as_accurate_as_fp32_fma(eng, use_tanh):
    rng = np.random.default_rng(0)
    W = _weights(3, gain3=40.0)                # the 0.01-gain head would hide hidden-layer errors: boost it
    n = 8192
    x = rng.normal(0, 1.5, (n, 18)).astype(np.float32)
    eps = rng.normal(0, 1, (n, 3)).astype(np.float32)
    actor = eng.GaussianActorKernel(use_tanh=use_tanh).load_state_dict(W)
    d_x, d_eps = torch.from_numpy(x).cuda(), torch.from_numpy(eps).cuda()
    out = {}
    for tc in (False, True):
        mean = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        a, lp = actor.sample(obs=d_x, eps_in=d_eps, mean_out=mean, tc=tc)
        torch.cuda.synchronize()
        out[tc] = (mean.cpu().numpy(), a.cpu().numpy(), lp.cpu().numpy())
    truth = _fp64_mean(W, x, use_tanh)
    err_fma = np.abs(out[False][0] - truth)
    err_tc = np.abs(out[True][0] - truth)
    print(f"use_tanh={use_tanh}: |mean - fp64| FFMA2 max {err_fma.max():.2e} rms {np.sqrt((err_fma**2).mean()):.2e}; "
          f"tensor cores max {err_tc.max():.2e} rms {np.sqrt((err_tc**2).mean()):.2e}")
    assert err_tc.max() < 2e-6                                           # fp32-level (plain bf16 / TF32 would be ~1e-3)
    # the bar that makes the tensor-core path the default: its rms error against the fp64 truth is no larger than the fp32
    # FFMA2 kernel's (measured 6.5e-8 vs 7.2e-8 with tanh, 3.3e-8 vs 3.5e-8 with ReLU); single worst elements within 1.25x
    assert np.sqrt((err_tc ** 2).mean()) <= np.sqrt((err_fma ** 2).mean())
    assert err_tc.max() <= 1.25 * err_fma.max() + 2e-8
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-4, atol=1e-4)


@pytest.mark.timeout(180)
def test_tc_ragged_sizes_and_env_state_path(eng):
    rng = np.random.default_rng(1)
    W = _weights(5, gain3=10.0)
    actor = eng.GaussianActorKernel().load_state_dict(W)
    for n in (1, 127, 128, 129, 1000):
        x = torch.from_numpy(rng.normal(0, 1, (n, 18)).astype(np.float32)).cuda()
        a0, l0 = actor.sample(obs=x, seed=7, step=3, tc=False)
        a1, l1 = actor.sample(obs=x, seed=7, step=3, tc=True)       # same Philox stream: same draws
        torch.testing.assert_close(a1, a0, rtol=0, atol=1e-5)
        torch.testing.assert_close(l1, l0, rtol=1e-4, atol=1e-4)
    # observation rebuilt and normalised from the fp64 env state inside the kernel
    env = eng.EnvBatch(777, mode="cw", d_capture=20000.0, max_episode_steps=10)
    P = np.array([200000.0, 0, 0]) + rng.normal(0, 3e4, (777, 3)); E = np.array([18000.0, 0, 0]) + rng.normal(0, 3e4, (777, 3))
    env.set_state(P, rng.normal(0, 3.0, (777, 3)), E, rng.normal(0, 3.0, (777, 3)))
    st = eng.RunningStats(18)
    st.update_normalize(env.observe())
    o0 = torch.empty((777, 18), dtype=torch.float32, device="cuda"); o1 = torch.empty_like(o0)
    a0, l0 = actor.sample(env=env, obs_stats=st, seed=1, step=2, obs_out=o0, tc=False)
    a1, l1 = actor.sample(env=env, obs_stats=st, seed=1, step=2, obs_out=o1, tc=True)
    assert torch.equal(o0, o1)
    torch.testing.assert_close(a1, a0, rtol=0, atol=1e-5)
    torch.testing.assert_close(l1, l0, rtol=1e-4, atol=1e-4)


@pytest.mark.timeout(180)
def test_observe_norm_is_the_fused_paths_arithmetic(eng):
    """sat_env_observe_norm (the fp32 observation the tensor-core actor path reads): (x - mean) / (std + 1e-8) in fp64, then the
    cast - environment.py:76-77 + normalization.py:41 - bit for bit, with and without statistics, ragged sizes"""
    import ctypes as C
    from ppo_rl_satellite_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(3)
    for n in (1, 127, 128, 1000):
        env = eng.EnvBatch(n, mode="cw", d_capture=20000.0, max_episode_steps=10)
        env.set_state(np.array([200000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)),
                      np.array([18000.0, 0, 0]) + rng.normal(0, 3e4, (n, 3)), rng.normal(0, 3.0, (n, 3)))
        st = eng.RunningStats(18)
        x = env.observe()                                       # fp64 [n, 18]
        st.update_normalize(x)
        out = torch.empty((n, 18), dtype=torch.float32, device="cuda")
        L.check(lib.sat_env_observe_norm(C.byref(env.st), L.ptr(st.buf), L.ptr(out), L.stream_ptr()), "observe_norm")
        ref = ((x - st.mean) / (st.std + 1e-8)).float()
        assert torch.equal(out, ref)
        L.check(lib.sat_env_observe_norm(C.byref(env.st), None, L.ptr(out), L.stream_ptr()), "observe_norm")
        assert torch.equal(out, x.float())
