"""CPU: the oracle port reproduces the REAL reference bit for bit.

(1) replaying the random draws recorded from the reference (tests/golden/*.npz),
(2) drawing from an identically seeded numpy Generator (same numpy => same stream),
(3) when /root/reference is present: stepping oracle and reference side by side.
"""
import hashlib

import numpy as np
import pytest

import oracle
from helpers import first, golden_names, load_golden, make_oracle, recorded_draws
from oracle.draws import GeneratorDraws
from oracle.ref_loader import reference_available

ALL = golden_names()


def _rollout(env, fx, check=True):
    h = hashlib.sha256()
    feed = lambda x: h.update(np.ascontiguousarray(x).astype(np.float64).tobytes())
    o = first(env.reset())
    feed(o)
    if check:
        np.testing.assert_array_equal(np.asarray(o, dtype=np.float64), fx["obs0"].astype(np.float64))
    for t in range(fx["meta"]["T"]):
        o, r, d, tr, _ = env.step(fx["actions"][t])
        feed(o); feed(r); feed(d); feed(tr)
        if check:
            np.testing.assert_array_equal(np.asarray(o, dtype=np.float64), fx["obs"][t].astype(np.float64), err_msg=f"obs t={t}")
            assert r.dtype == np.float32
            np.testing.assert_array_equal(r, fx["rew"][t], err_msg=f"rew t={t}")
            np.testing.assert_array_equal(d, fx["term"][t], err_msg=f"term t={t}")
            np.testing.assert_array_equal(tr, fx["trunc"][t], err_msg=f"trunc t={t}")
    return h.hexdigest()


@pytest.mark.parametrize("name", ALL)
def test_oracle_replays_reference_draws(name):
    fx = load_golden(name)
    draws = recorded_draws(fx)
    env = make_oracle(fx["meta"], draws)
    sha = _rollout(env, fx)
    assert draws.exhausted, "oracle consumed fewer random draws than the reference"
    assert sha == fx["meta"]["sha256"]
    st = env.state
    if "state_heavens" in fx:
        np.testing.assert_array_equal(st["s"], fx["state_s"])
        np.testing.assert_array_equal(st["heavens"], fx["state_heavens"])
        np.testing.assert_array_equal(st["priests"], fx["state_priests"])
    elif "state_s" in fx:
        np.testing.assert_array_equal(st["s"], fx["state_s"])
        np.testing.assert_array_equal(st["ndrop"], fx["state_ndrop"].astype(np.int64))
    else:
        np.testing.assert_array_equal(st["agent"], fx["state_agent"])
        np.testing.assert_array_equal(st["goal"], fx["state_goal"])
        if "state_velocity" in fx:
            np.testing.assert_array_equal(st["velocity"], fx["state_velocity"])
    np.testing.assert_array_equal(st["elapsed"], fx["state_elapsed"])


@pytest.mark.parametrize("name", ALL)
def test_oracle_same_seed_same_stream(name):
    fx = load_golden(name)
    if fx["meta"]["numpy"] != np.__version__:
        pytest.skip("numpy version differs from the one that generated the fixture")
    env = make_oracle(fx["meta"], GeneratorDraws(seed=fx["meta"]["seed"]))
    assert _rollout(env, fx, check=False) == fx["meta"]["sha256"]


# hashes taken at survey time from the reference
APPENDIX_D = {
    "taxi": "fb7ba036d23923d1", "taxi_hansen": "1bb378c58a882176", "taxi_ext": "856ff1f13ca57ed4",
    "taxi_ext_hansen": "2c41d2708eb3628d", "taxi_multi": "b70046e40f4fc21e", "rooms_hansen8": "dc1a7f8b12eadc90",
    "rooms_hansen4_card": "39fb87df506c4bb5", "rooms_vhansen8": "0097d0fbcd83bd98", "rooms_vghansen8": "96f326d4a3b84a4a",
    "rooms_grid5": "d70b2b6999e79baf", "rooms_grid9": "2289093732703d13", "rooms32_grid9_rgoal": "52a77d8b202725f0",
    "rooms_mdp": "f3babe8ade837b4d", "crooms_vmdp": "7725fcf136b2d644", "crooms_hansen8_ord": "a2fae1ca1ab3ce71",
}


@pytest.mark.parametrize("name", sorted(APPENDIX_D))
def test_fixture_matches_survey_hash(name):
    assert load_golden(name)["meta"]["sha256"].startswith(APPENDIX_D[name])


def test_layout_asset_matches_reference_grids():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "layout_grids.npz"))
    assert set(oracle.LAYOUT_NAMES) == {k[5:] for k in z.files}
    for name in oracle.LAYOUT_NAMES:
        np.testing.assert_array_equal(oracle.load_layout(name), z["grid_" + name].astype(np.int64))


def test_tag_move_target_matches_reference_vectors():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "tag_move_target.npz"))
    out = oracle.tag_move_target(z["ant"], z["target"], z["choice"])
    # The scalar reference normalises with np.linalg.norm(v) == sqrt(BLAS dot(v, v)), whose FMA
    # contraction differs by <= 1 ulp from the vectorized sqrt(sum(v*v)); tolerance: 4 ulp (1e-15 rel).
    np.testing.assert_allclose(out, z["new_target"], rtol=1e-15, atol=1e-15)
    assert np.array_equal(out == z["target"], z["new_target"] == z["target"])  # stay/blocked decisions identical


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", ["taxi", "taxi_multi", "rooms_hansen8", "rooms32_grid9_rgoal", "crooms_vel_rg",
                                  "msrooms_hansen8_3floors", "msrooms_vghansen_rg"])
def test_oracle_lockstep_with_live_reference(name):
    from oracle.draws import make_generator
    from oracle.ref_loader import load_reference
    E = load_reference()
    fx = load_golden(name)
    meta = fx["meta"]
    kw = dict(meta["kwargs"])
    ref = getattr(E, meta["cls"])(256, **kw)
    if meta["cls"] == "CRoomsEnv":
        ref.rng = make_generator(123)
    else:
        ref._np_random = make_generator(123)
    orc = make_oracle(meta, GeneratorDraws(seed=123), num_envs=256)
    np.testing.assert_array_equal(first(ref.reset()), first(orc.reset()))
    arng = np.random.default_rng(5)
    for t in range(300):
        a = arng.uniform(-1, 1, (256, 2)) if meta["n_act"] == 0 else arng.integers(meta["n_act"], size=256)
        ro, oo = ref.step(a.copy()), orc.step(a.copy())
        for x, y in zip(ro[:4], oo[:4]):
            np.testing.assert_array_equal(x, y, err_msg=f"t={t}")
