"""Run under torchrun with any world size: prints statistics that must be IDENTICAL (bit for bit) for every world size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_fishing_b200 as gf
from gym_fishing_b200.distributed import init_from_env
rank, world, local = init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
S = float.fromhex("0x1.9374bcp-1")
a = gf.rollout_stats("fishing-v2", 1 << 20, 300, policy="escapement", S=S, device=dev, sigma=0.05, Tmax=1000)
b = gf.escapement_sweep("fishing-v1", levels=256, replicates=1024, T=200, device=dev, sigma=0.1, Tmax=1000)
if rank == 0:
    print("INVARIANT", world, float(a["sum"]).hex(), float(a["sumsq"]).hex(), float(a["n_done"]),
          float(b["sum"].sum()).hex(), float(b["sumsq"][17]).hex(), float(b["mean"][200]).hex())
if world > 1:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
