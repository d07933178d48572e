"""Small run of every kernel family for compute-sanitizer (memcheck): Philox + replay, ragged sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gym-po-taxi_b200"))
import numpy as np, torch
import oracle
from gym_po.envs import TaxiVecEnv, RoomsEnv, CRoomsEnv, TagVecEnv, MultistoryFourRoomsEnv, CarVecEnv
from gym_po.wrappers import NormalizeReward, RecordEpisodeStatistics
dev = "cuda:0"
for b in (1, 700, 5000):
    for hansen in (False, True):
        e = TaxiVecEnv(b, hansen_obs=hansen, time_limit=7, device=dev, seed=1, track_stats=hansen)
        e.reset()
        for _ in range(25):
            e.step(torch.randint(0, 5, (b,), dtype=torch.int8, device=dev))
        e.step_host(np.zeros(b, dtype=np.int8))
    for obs, n, goal in (("hansen8", 3, (0, 0)), ("vector_goal_hansen8", 3, None), ("grid", 5, None), ("grid", 9, (0, 0)), ("grid", 4, None), ("mdp_goal", 3, None)):
        for layout in ("4", "32"):
            e = RoomsEnv(b, layout, obs_type=obs, obs_n=n, goal_xy=goal, time_limit=6, device=dev, seed=2, track_stats=(obs == "grid"))
            e.reset()
            for _ in range(20):
                e.step(torch.randint(0, 8, (b,), dtype=torch.int8, device=dev))
    for prec in ("float64", "float32"):
        e = CRoomsEnv(b, "8", obs_type="grid", obs_m=5, goal_xy=None, use_velocity=True, time_limit=6, device=dev, seed=3, precision=prec)
        e.reset()
        for _ in range(20):
            e.step(torch.rand((b, 2), device=dev) * 2 - 1)
        e = TagVecEnv(b, time_limit=6, device=dev, seed=4, precision=prec)
        e.reset()
        for _ in range(20):
            e.step(torch.rand((b, 2), device=dev) * 2 - 1)
    # multistory rooms (all observation byte widths), car-flag, wrapper layer, fused multi-step launches
    for obs, goal in (("mdp", (9, 7, -1)), ("vector_mdp", (9, 7, -1)), ("vector_mdp_goal", None), ("vector_goal_hansen8", None), ("hansen", None)):
        e = MultistoryFourRoomsEnv(b, grid_z=3, obs_type=obs, goal_xyz=goal, time_limit=6, device=dev, seed=5)
        e.reset()
        for _ in range(20):
            e.step(torch.randint(0, 4, (b,), dtype=torch.int8, device=dev))
        e.step_many(torch.randint(0, 4, (9, e.capacity), dtype=torch.int8, device=dev))
    e = CarVecEnv(b, time_limit=6, device=dev, seed=6)
    e.reset()
    for _ in range(20):
        e.step(torch.rand((b, 1), device=dev) * 2 - 1)
    w = NormalizeReward(RecordEpisodeStatistics(TaxiVecEnv(b, time_limit=7, device=dev, seed=7)), gamma=0.9)
    w.reset()
    for _ in range(20):
        w.step(torch.randint(0, 5, (b,), dtype=torch.int8, device=dev))
    for env, n_act in ((TaxiVecEnv(b, time_limit=7, num_passengers=2, device=dev, seed=8), 5),
                       (RoomsEnv(b, "8", obs_type="grid", obs_n=7, goal_xy=None, time_limit=6, device=dev, seed=9), 8)):
        env.reset()
        env.step_many(torch.randint(0, n_act, (11, env.capacity), dtype=torch.int8, device=dev))
# replay mode, arith kernel
big = ("A" + " " * 13 + "B",) + (" " * 6 + "|" + " " * 8,) * 6 + ("C" + " " * 13 + "D",) + (" " * 15,) * 6 + ("E" + " " * 13 + "F",)
orc = oracle.TaxiOracle(600, map=big, time_limit=5, draws=oracle.GeneratorDraws(seed=1))
e = TaxiVecEnv(600, map=big, time_limit=5, device=dev, rng_mode="replay")
orc.reset(); e.set_replay(**orc.draws); e.reset()
for _ in range(15):
    a = np.random.default_rng(0).integers(5, size=600)
    orc.step(a); e.set_replay(**orc.draws); e.step(torch.as_tensor(a, dtype=torch.int8, device=dev))
torch.cuda.synchronize()
print("sanitize smoke done")
