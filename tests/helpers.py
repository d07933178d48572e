"""Shared helpers for the tests: golden-fixture loading and oracle construction."""
import glob
import json
import os

import numpy as np

import oracle
from oracle.draws import RecordedDraws

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = ["multinomial_argmax", "integers", "random", "choice", "normal", "uniform"]
INT_KINDS = {"multinomial_argmax", "integers", "choice"}


def golden_names(prefix=""):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if n.startswith(prefix) and n not in ("tag_move_target", "layout_grids") and not n.startswith("render_")]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fx = {k: z[k] for k in z.files}
    fx["meta"] = json.loads(str(fx["meta"]))
    return fx


def recorded_draws(fx):
    log, pos = [], 0
    for k, n in zip(fx["log_kind"], fx["log_size"]):
        kind = KINDS[int(k)]
        v = fx["log_val"][pos:pos + int(n)]
        pos += int(n)
        integral = kind in INT_KINDS and np.array_equal(v, np.round(v))   # choice([-0.5, 0.5]) stays float
        log.append((kind, v.astype(np.int64) if integral else v))
    return RecordedDraws(log)


_TAXI_ALIASES = {
    "TaxiVecEnv": {},
    "HansenTaxiVecEnv": {"hansen_obs": True},
    "ExtendedTaxiVecEnv": {"map": oracle.EXTENDED_TAXI_MAP},
    "ExtendedHansenTaxiVecEnv": {"hansen_obs": True, "map": oracle.EXTENDED_TAXI_MAP},
}


def make_oracle(meta, draws=None, num_envs=None):
    """Oracle env for a fixture's (class, kwargs)."""
    cls, kw = meta["cls"], dict(meta["kwargs"])
    b = num_envs or meta["B"]
    if "goal_xy" in kw and kw["goal_xy"] is not None:
        kw["goal_xy"] = tuple(kw["goal_xy"])
    if cls in _TAXI_ALIASES:
        kw.update(_TAXI_ALIASES[cls])
        return oracle.TaxiOracle(b, draws=draws, **kw)
    if cls == "RoomsEnv":
        return oracle.RoomsOracle(b, draws=draws, **kw)
    if cls == "CRoomsEnv":
        return oracle.CRoomsOracle(b, draws=draws, **kw)
    if cls == "MultistoryFourRoomsEnv":
        if "goal_xyz" in kw and kw["goal_xyz"] is not None:
            kw["goal_xyz"] = tuple(kw["goal_xyz"])
        return oracle.MSRoomsOracle(b, draws=draws, **kw)
    if cls in ("CarVecEnv", "DiscreteActionCarVecEnv"):
        return oracle.CarOracle(b, draws=draws, **kw)
    raise KeyError(cls)


def first(out):
    return out[0] if isinstance(out, tuple) else out
