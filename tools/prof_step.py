import sys, cProfile, pstats, torch
sys.path.insert(0, "/root/repo")
import gym_fishing_b200 as gf
n = 4096
v = gf.FishingVecEnv("fishing-v1", num_envs=n, device="cuda", io="torch", seed=0, sigma=0.01)
v.reset()
a = torch.rand(n, 1, dtype=torch.float64, device="cuda") * 2 - 1
for _ in range(500): v.step(a)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20000): v.step(a)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(18)
