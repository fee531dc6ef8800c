import sys, os
sys.path.insert(0, os.getcwd())
import torch
from probabilisticsemslam_b200 import api, synth, device as dev
api.set_murty_path("cta")
pb = synth.g1_dense(1, nM=5)
plan = dev.MurtyPlan(pb, k=200, weights=True)
for rep in range(3): plan.run()
torch.cuda.synchronize()
