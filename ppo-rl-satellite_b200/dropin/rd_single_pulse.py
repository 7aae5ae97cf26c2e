"""rd_single_pulse -- drop-in for the sweep of single_pluse_model/RD_single_pulse.py (:9-148).

`params` is the module-level dict of the reference (:9-20); `Incoming_parameters(data, delta_max)` stores the six
elements like :22-34 and runs the sweep. The reference then fits an ellipse to the two point clouds with sklearn
(curve_fitting.py) and plots them; that post-processing is not part of this library, so `Reachable_Domain()` returns the
point clouds (RF_max, RF_min) the reference hands to `cf.Curve_fitting` (:146) - a maintainer keeps the upstream fit and
replaces only the 201 x 201 x 2 fsolve loop. `reachable_domain_batch` is the batched form (many states per launch).
"""
import numpy as np

try:
    from ._boot import engine as _eng
except ImportError:
    from _boot import engine as _eng

params = {"a": 10 ** 7, "i": 0, "e0": 0.2, "f": np.pi / 2, "delta_max": 500, "u": 3.986e14, "N1": 1, "N2": 200, "N3": 200,
          "delta_l": 1500}


def reachable_domain_batch(elements, delta_max, N2=None, N3=None, u=None):
    """elements [n,6] (a, e, i, omega, Omega, f), delta_max [n] -> list of (RF_max [m,3], RF_min [m,3]) per state."""
    import torch
    el = np.ascontiguousarray(np.asarray(elements, dtype=np.float64).reshape(-1, 6))
    dm = np.broadcast_to(np.asarray(delta_max, dtype=np.float64), (el.shape[0],)).copy()
    hi, lo, valid = _eng.reachable_domain(torch.from_numpy(el).cuda(), torch.from_numpy(dm).cuda(),
                                          params["N2"] if N2 is None else N2, params["N3"] if N3 is None else N3,
                                          params["u"] if u is None else u)
    hi, lo, valid = hi.cpu().numpy(), lo.cpu().numpy(), valid.cpu().numpy().astype(bool)
    return [(hi[k][valid[k]], lo[k][valid[k]]) for k in range(el.shape[0])]


def Reachable_Domain():
    if params["N1"] != 1:
        raise NotImplementedError("the reference sweeps N1 = 1 (:16, :63); other values are not supported")
    el = [params["a"], params["e0"], params["i"], 0.0, 0.0, params["f"]]
    return reachable_domain_batch([el], [params["delta_max"]])[0]


def Incoming_parameters(data, delta_max):
    params["a"], params["i"], params["e0"], params["f"] = data[0], data[2], data[1], data[5]
    params["delta_max"] = delta_max
    return Reachable_Domain()
