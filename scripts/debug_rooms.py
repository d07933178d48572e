import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gym-po-taxi_b200"))
import torch
from gym_po.envs import RoomsEnv
for layout in ("4", "32"):
    for obs, n in (("mdp", 3), ("hansen8", 3), ("vector_hansen8", 3), ("grid", 3), ("grid", 5), ("grid", 9), ("grid", 11)):
        for goal in ((0, 0), None):
            try:
                e = RoomsEnv(64, layout, obs_type=obs, obs_n=n, goal_xy=goal, seed=0)
                e.reset()
                e.step(torch.zeros(e.capacity, dtype=torch.int8, device="cuda"))
                torch.cuda.synchronize()
                print("ok  ", layout, obs, n, goal)
            except Exception as ex:
                print("FAIL", layout, obs, n, goal, str(ex)[:150])
