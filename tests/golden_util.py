import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kitti07_pair.npz")


def unpack_cov6(c6):
    c = np.zeros((len(c6), 3, 3))
    c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2] = c6.T
    c[:, 1, 0], c[:, 2, 0], c[:, 2, 1] = c6[:, 1], c6[:, 2], c6[:, 4]
    return c


def load():
    z = np.load(GOLDEN)
    d = {k: z[k] for k in z.files}
    d["target_points"] = d["target_points"].astype(np.float64)
    d["source_points"] = d["source_points"].astype(np.float64)
    d["target_covs"] = unpack_cov6(d["target_covs"])
    d["source_covs"] = unpack_cov6(d["source_covs"])
    d["resolution"] = float(d["resolution"])
    return d
