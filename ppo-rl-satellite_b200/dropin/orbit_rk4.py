"""orbit_rk4 -- drop-in for the functions of the reference's RK4 script "轨道外推-龙格库塔算法.py" (:9-40).

StateEq(t, RV) and RungeKutta(t0, r0, h) accept a 6-vector or a (6, N) array exactly like the script's functions
(which broadcast), and run the CUDA propagator. The module-level constants mu, Re, J2 are read at call time, so
`orbit_rk4.J2 = 0.0` gives pure two-body motion, as editing the script's globals does.
"""
import numpy as np

try:
    from ._boot import engine as _eng, _lib as _L
except ImportError:  # imported as a top-level module (dropin/ on sys.path, the CPPO_main.py case)
    from _boot import engine as _eng, _lib as _L

mu = 398600
Re = 6378.137
J2 = 0.00108263


def _to_dev(RV):
    import torch
    a = np.asarray(RV, dtype=np.float64)
    one = a.ndim == 1
    a2 = a.reshape(6, -1)
    x, _buf = _eng.alloc_soa(6, a2.shape[1], torch.float64, "cuda")
    x.copy_(torch.from_numpy(np.ascontiguousarray(a2)))
    return x, one


def StateEq(t, RV):
    import torch
    x, one = _to_dev(RV)
    f, _buf = _eng.alloc_soa(6, x.shape[1], torch.float64, "cuda")
    _L.check(_L.load().sat_state_eq(x.data_ptr(), f.data_ptr(), x.shape[1], x.stride(0), float(mu), float(Re), float(J2),
                                    _L.stream_ptr()), "sat_state_eq")
    out = f.cpu().numpy()
    return out[:, 0].copy() if one else out


def RungeKutta(t0, r0, h, steps=1):
    x, one = _to_dev(r0)
    _eng.rk4_propagate(x, float(h), int(steps), mu=float(mu), re=float(Re), j2=float(J2))
    out = x.cpu().numpy()
    return out[:, 0].copy() if one else out
