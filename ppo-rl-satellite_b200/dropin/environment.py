"""environment.satellites -- drop-in for the reference's environment.py, backed by the fused CUDA env step.

Same constructor, attributes and reset/step contract as the reference (environment.py:26-79, 81-255):
    env = satellites(..., d_capture=50000, args=args)
    s = env.reset(Flag)                       # ndarray[18] (int64, like the reference's integer reset arrays)
    s_, r, done = env.step(pursuer_action, escaper_action, epsiode_count)
One environment = a batch of size 1 on the GPU (cw mode: the propagation the shipped env uses). The batched
engine lives in ppo_rl_satellite_b200.engine.EnvBatch; there is no CPU path.
Not replicated: the per-step print (environment.py:134) and, under Flag == 2, the surrogate-network fit (:295-301; the
step's dynamics themselves are).
"""
import numpy as np

try:
    from ._boot import engine as _eng
except ImportError:  # imported as a top-level module (dropin/ on sys.path, the CPPO_main.py case)
    from _boot import engine as _eng

try:  # the reference imports gym for two attribute objects only (environment.py:57,62)
    from gym import spaces as _spaces
    _Box, _Discrete = _spaces.Box, _spaces.Discrete
except Exception:  # gym is optional
    class _Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class _Discrete:
        def __init__(self, n):
            self.n, self.shape = n, ()


class satellites:
    __annotations__ = {"Pursuer_position": np.ndarray, "Pursuer_vector": np.ndarray,
                       "Escaper_position": np.ndarray, "Escaper_vector": np.ndarray}

    def __init__(self, Pursuer_position=np.array([2000, 2000, 1000]), Pursuer_vector=np.array([1.71, 1.14, 1.3]),
                 Escaper_position=np.array([1000, 2000, 0]), Escaper_vector=np.array([1.71, 1.14, 1.3]),
                 M=0.4, dis_safe=1000, d_capture=100000, Flag=0, fuel_c=320, fuel_t=320, d_range=100000, args=None):
        import torch
        self._torch = torch
        # constructor positions are dead in the reference too: reset() overwrites them
        self.dis_dafe, self.M, self.Flag = dis_safe, M, Flag
        self.burn_reward, self.win_reward = 0, 100
        self.max_episode_steps = args.max_episode_steps
        self.ellipse_params = []
        self._env = _eng.EnvBatch(1, mode="cw", flag=Flag if Flag in (0, 1, 2) else 0, d_capture=d_capture,
                                  d_range=d_range, fuel_c=fuel_c, fuel_t=fuel_t,
                                  max_episode_steps=self.max_episode_steps, auto_reset=False)
        self._cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._obs = torch.empty((1, 18), dtype=torch.float64, device="cuda")
        self._pa = torch.empty((1, 3), dtype=torch.float64, device="cuda")
        self._ea = torch.empty((1, 3), dtype=torch.float64, device="cuda")
        position_low = np.array([-500000] * 3 + [-10000000] * 6)
        velocity_low = np.array([-10000] * 3 + [-50000] * 6)
        self.observation_space = _Box(low=np.concatenate((position_low, velocity_low)),
                                      high=-np.concatenate((position_low, velocity_low)), shape=(18,), dtype=np.float32)
        self.action_space = np.array([[-1.6, 1.6], [-1.6, 1.6], [-1.6, 1.6]])
        self.action_space_beta = _Discrete(5)
        self.pursuer_reward = 0.0
        self.escaper_reward = 0.0

    # ---- attributes the reference exposes, read from the device state
    d_capture = property(lambda self: self._env.params.d_capture,
                         lambda self, v: setattr(self._env.params, "d_capture", float(v)))
    d_range = property(lambda self: self._env.params.d_range,
                       lambda self, v: setattr(self._env.params, "d_range", float(v)))
    Pursuer_position = property(lambda self: self._env.state[0:3, 0].cpu().numpy())
    Pursuer_vector = property(lambda self: self._env.state[3:6, 0].cpu().numpy())
    Escaper_position = property(lambda self: self._env.state[6:9, 0].cpu().numpy())
    Escaper_vector = property(lambda self: self._env.state[9:12, 0].cpu().numpy())
    fuel_c = property(lambda self: float(self._env.fuel_c[0]))
    fuel_t = property(lambda self: float(self._env.fuel_t[0]))
    dis = property(lambda self: float(self._env.dis[0]))
    dangerous_zone = property(lambda self: int(self._env.dangerous_zone[0]))

    def reset(self, Flag):
        if Flag not in (0, 1, 2):
            raise ValueError("Flag must be 0, 1 or 2")
        # Flag 2 (environment.py:257-316): the step's dynamics (gating, impulse, fuel, CW propagation, terminal checks, reward 0)
        # run on the device; the reachable-domain surrogate fit the reference interleaves (:295-301) is outside the hot path
        self.Flag = Flag
        self._env.reset(flag=Flag)
        self.pursuer_reward = 0.0
        self.escaper_reward = 0.0
        return self._env.observe().cpu().numpy()[0].astype(np.int64)      # integer arrays, environment.py:67-77

    def step(self, pursuer_action, escaper_action, epsiode_count):
        t = self._torch
        self._pa.copy_(t.as_tensor(np.asarray(pursuer_action, dtype=np.float64).reshape(1, 3)))
        self._ea.copy_(t.as_tensor(np.asarray(escaper_action, dtype=np.float64).reshape(1, 3)))
        self._cnt.fill_(int(epsiode_count))
        r, d = self._env.step(self._pa, self._ea, count=self._cnt, obs_f64=self._obs)
        out = t.cat([self._obs.reshape(-1), r.reshape(-1), d.reshape(-1).double()]).cpu().numpy()
        reward, done = float(out[18]), bool(out[19])
        if self.Flag == 0:
            self.pursuer_reward = reward
        elif self.Flag == 2:
            self.pursuer_reward = 0
        else:
            self.escaper_reward, self.pursuer_reward = reward, -reward
        return out[:18].copy(), reward, done

    def calculate_number_hanger_area(self):
        """environment.py:317-332 on the current state (also refreshed by every step())."""
        t = self._torch
        s = self._env.state[:, 0]
        Rcw = t.tensor([27098000.0, 32306000.0, 0.0], dtype=t.float64, device="cuda")
        Vcw = t.tensor([-2350.0, 1970.0, 0.0], dtype=t.float64, device="cuda")
        rv = t.cat([Rcw + s[0:3], Vcw + s[3:6], Rcw + s[6:9], Vcw + s[9:12]]).reshape(1, 12)
        dz = int(_eng.danger_zone_count(rv, s[12:13].clone())[0])
        self._env.dangerous_zone[0] = dz
        return dz

    @staticmethod
    def relative_state_to_absolute_state(R0, V0):
        assert isinstance(R0, np.ndarray) and isinstance(V0, np.ndarray)
        return np.array([27098000, 32306000, 0]) + R0, np.array([-2350, 1970, 0]) + V0   # environment.py:338-341
