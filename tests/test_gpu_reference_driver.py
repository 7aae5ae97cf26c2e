"""BASELINE config 1: the reference's own, UNMODIFIED CPPO_main.py drives the drop-in modules.

baseline/_ref/ (git-ignored, staged by tools/stage_reference.py in the build container; it travels to the GPU box with the
snapshot) holds the upstream files. dropin/ is put ahead of it on sys.path, so `from environment import satellites`,
`from ppo_continuous import PPO_continuous`, `from replaybuffer import ReplayBuffer`, `from normalization import ...`
inside CPPO_main.py resolve to the CUDA-backed drop-ins while CPPO_main.py itself and plot_function.py are the reference's
files (CPPO_main.py:1-9). The three packages the image lacks (gym, matplotlib, plotly) are stubbed as in SURVEY.md s8c.
Skipped when the directory is absent."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "ppo-rl-satellite_b200", "dropin")
REFDIR = os.path.join(ROOT, "baseline", "_ref")
NAMES = ("environment", "satellite_function", "ppo_continuous", "replaybuffer", "normalization", "CPPO_main", "plot_function")


@pytest.fixture(scope="module")
def cppo_main():
    if not os.path.isfile(os.path.join(REFDIR, "CPPO_main.py")):
        pytest.skip("baseline/_ref not staged (python tools/stage_reference.py in the build container)")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refshim
    refshim._install_stubs()
    saved = {k: sys.modules.pop(k, None) for k in NAMES}
    sys.path.insert(0, REFDIR)
    sys.path.insert(0, DROPIN)               # drop-ins win over the reference's same-named modules
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import CPPO_main
        import environment, ppo_continuous, replaybuffer, normalization
        # the driver is the reference's file, everything it imports for the hot path is ours
        assert os.path.samefile(CPPO_main.__file__, os.path.join(REFDIR, "CPPO_main.py"))
        for m in (environment, ppo_continuous, replaybuffer, normalization):
            assert os.path.dirname(os.path.abspath(m.__file__)) == DROPIN, m.__file__
        assert CPPO_main.satellites is environment.satellites and CPPO_main.PPO_continuous is ppo_continuous.PPO_continuous
        yield CPPO_main
    finally:
        sys.path.remove(DROPIN); sys.path.remove(REFDIR)
        for k in NAMES:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]


def _env(M, args):
    # CPPO_main.py:330-335
    return M.satellites(Pursuer_position=np.array([2000000, 2000000, 1000000]), Pursuer_vector=np.array([1710, 1140, 1300]),
                        Escaper_position=np.array([1850000, 2000000, 1000000]), Escaper_vector=np.array([1710, 1140, 1300]),
                        d_capture=50000, args=args)


def test_unmodified_cppo_main_trains_and_tests_on_the_dropins(cppo_main, tmp_path):
    """train_pursuer_network (CPPO_main.py:94-161) for 4 episodes with updates, then test_network (:233-282) from the
    checkpoint it saved: the reference's own functions, called with the reference's own __main__ arguments (:327-335)."""
    M = cppo_main
    with contextlib.redirect_stdout(io.StringIO()):
        args = M.args_param(max_episode_steps=64, batch_size=64, max_train_steps=4, K_epochs=3, chkpt_dir=str(tmp_path))
    env = _env(M, args)
    torch.manual_seed(0); np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        agent = M.train_pursuer_network(args, env, show_picture=False, pre_train=False, d_capture=15000)
    assert env.d_capture == 15000
    assert os.path.exists(os.path.join(str(tmp_path), "agent_pursuer_actor_Gaussian"))
    assert os.path.exists(os.path.join(str(tmp_path), "agent_pursuer_critic"))
    assert all(torch.isfinite(p).all() for p in agent.actor.parameters())
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        M.test_network(args, env, show_pictures=False, d_capture=20000)
    assert "当前测试得分为" in out.getvalue()          # CPPO_main.py:278 printed the episode score


def test_unmodified_cppo_main_loop_reproduces_the_reference_stream(cppo_main, golden, tmp_path):
    """The same train_pursuer_network loop with the agents' actions replaced by the action stream recorded from the reference
    run (fixture env_golden.npz::cfg1_*, 10 episodes of 64 steps): every (obs, reward, done) the drop-in env hands back to
    CPPO_main.py equals the reference env's, bit for bit, and the PPO update runs every 64 stored transitions."""
    M = cppo_main
    g = golden("env_golden.npz")
    import ppo_continuous as drop_ppo
    t = {"i": 0, "updates": 0}

    class Replay(drop_ppo.PPO_continuous):
        def __init__(self, a, idx):
            super().__init__(a, idx)
            self.key = "cfg1_pa" if idx == "pursuer" else "cfg1_ea"

        def choose_action(self, s):
            _, logp = super().choose_action(s)                    # the real sampling kernel runs; its draw is replaced
            return g[self.key][t["i"]], logp

        def update(self, rb, total_steps):
            t["updates"] += 1
            return super().update(rb, total_steps)

    with contextlib.redirect_stdout(io.StringIO()):
        args = M.args_param(max_episode_steps=64, batch_size=64, max_train_steps=10, K_epochs=3, chkpt_dir=str(tmp_path))
    env = _env(M, args)
    stream = []
    real_step = env.step

    def step(pa, ea, count):
        out = real_step(pa, ea, count)
        stream.append(out)
        t["i"] += 1
        return out
    env.step = step
    saved = M.PPO_continuous
    M.PPO_continuous = Replay
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            M.train_pursuer_network(args, env, show_picture=False, pre_train=False, d_capture=20000)
    finally:
        M.PPO_continuous = saved
    assert len(stream) == 640 and t["updates"] == 10
    obs = np.array([s[0] for s in stream]); rew = np.array([s[1] for s in stream]); done = np.array([s[2] for s in stream])
    assert np.array_equal(obs, g["cfg1_obs"]) and np.array_equal(rew, g["cfg1_reward"]) and np.array_equal(done, g["cfg1_done"])
