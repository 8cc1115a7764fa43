import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import gtsam_points_b200 as g
from gtsam_points_b200 import synthetic as syn
tp, tc = syn.make_cloud(40000, stream=1, scale=0.25)
sp, sc = syn.make_cloud(20000, stream=2, scale=0.25)
delta = syn.random_pose(np.random.default_rng(1), 0.02, 0.2)
values = {0: np.eye(4), 1: delta}
vm = g.GaussianVoxelMapGPU(0.5)
vm.insert(g.PointCloud(tp, tc))
src = g.PointCloud(sp, sc)
f = g.IntegratedVGICPFactor(0, 1, vm, src)
f.linearize(values)
print("linearize ok", f.num_inliers())
print("error", f.error(values))
