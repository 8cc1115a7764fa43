"""A minimal Levenberg-Marquardt over Pose3 variables (test infrastructure).

Stands in for gtsam_points::LevenbergMarquardtOptimizerExt (src/gtsam_points/optimizers/levenberg_marquardt_ext.cpp:354-434),
which lives on top of GTSAM and is NOT part of this repository: one outer iteration = linearize every factor, then try
damped steps `(G + lambda I) d = g`, retract `T <- T Exp(d)` (GTSAM Pose3 tangent [rot, trans]) and evaluate error() with the
correspondences frozen at the linearization point, exactly the call sequence the reference optimizer drives.
Works with any factor exposing keys() / linearize(values) -> HessianFactor / error(values).
"""
import numpy as np

from gtsam_points_b200 import synthetic as syn
from gtsam_points_b200.factors import HessianFactor, pose_inverse


class OracleFactorAdapter:
    """Gives tests/oracle_lib.Factor the factor interface (keys / linearize / error on `values`)."""

    def __init__(self, ofactor, target_key, source_key, fixed_target_pose=None):
        self.f = ofactor
        self.fixed = fixed_target_pose
        self._keys = (source_key,) if fixed_target_pose is not None else (target_key, source_key)

    def keys(self):
        return self._keys

    def calc_delta(self, values):
        if self.fixed is not None:
            return pose_inverse(self.fixed) @ values[self._keys[0]]
        return pose_inverse(values[self._keys[0]]) @ values[self._keys[1]]

    def linearize(self, values):
        l = self.f.linearize(self.calc_delta(values))
        if self.fixed is not None:
            return HessianFactor(self._keys, None, None, None, l["H_source"], -l["b_source"], l["error"])
        return HessianFactor(self._keys, l["H_target"], l["H_target_source"], -l["b_target"], l["H_source"], -l["b_source"], l["error"])

    def error(self, values):
        return self.f.error(self.calc_delta(values))


class PriorFactor:
    """0.5 * || Log(T_prior^-1 T) ||^2_Sigma with a first-order Jacobian (identity): enough to pin a gauge pose."""

    def __init__(self, key, pose, precision=1e6):
        self.key, self.pose, self.precision = key, np.array(pose), precision

    def keys(self):
        return (self.key,)

    def _residual(self, values):
        D = pose_inverse(self.pose) @ values[self.key]
        w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
        return np.concatenate([w, D[:3, :3].T @ D[:3, 3]])

    def linearize(self, values):
        r = self._residual(values)
        H = self.precision * np.eye(6)
        return HessianFactor((self.key,), None, None, None, H, -self.precision * r, self.precision * float(r @ r))

    def error(self, values):
        r = self._residual(values)
        return self.precision * float(r @ r)


def optimize(factors, values, max_iterations=30, rel_tol=1e-6, abs_tol=1e-10, lambda_init=1e-5, on_iteration=None):
    values = {k: np.array(v, dtype=np.float64) for k, v in values.items()}
    keys = sorted(values.keys())
    index = {k: i for i, k in enumerate(keys)}
    n = 6 * len(keys)
    lam = lambda_init
    history = []
    for it in range(max_iterations):
        G = np.zeros((n, n))
        g = np.zeros(n)
        cur = 0.0
        for f in factors:
            hf = f.linearize(values)
            cur += hf.f
            ks = hf.keys
            if len(ks) == 2:
                a, b = 6 * index[ks[0]], 6 * index[ks[1]]
                G[a : a + 6, a : a + 6] += hf.H_target
                G[a : a + 6, b : b + 6] += hf.H_target_source
                G[b : b + 6, a : a + 6] += hf.H_target_source.T
                G[b : b + 6, b : b + 6] += hf.H_source
                g[a : a + 6] += hf.g_target
                g[b : b + 6] += hf.g_source
            else:
                b = 6 * index[ks[0]]
                G[b : b + 6, b : b + 6] += hf.H_source
                g[b : b + 6] += hf.g_source
        accepted = False
        for _ in range(12):
            try:
                d = np.linalg.solve(G + lam * np.eye(n), g)
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            trial = {k: values[k] @ syn.se3_exp(d[6 * index[k] : 6 * index[k] + 6]) for k in keys}
            new = sum(f.error(trial) for f in factors)
            if new < cur:
                values, accepted = trial, True
                lam = max(lam / 10, 1e-12)
                break
            lam *= 10
        history.append(dict(iteration=it, error=cur, new_error=new if accepted else cur, lam=lam, accepted=accepted))
        if on_iteration:
            on_iteration(history[-1], values)
        if not accepted:
            break
        if cur - new < abs_tol or (cur - new) / max(cur, 1e-300) < rel_tol:
            break
    return values, history


def pose_error(T_est, T_gt):
    D = pose_inverse(T_gt) @ T_est
    c = np.clip((np.trace(D[:3, :3]) - 1) / 2, -1, 1)
    return float(np.arccos(c)), float(np.linalg.norm(D[:3, 3]))
