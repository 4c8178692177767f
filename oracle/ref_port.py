"""CPU port of the reference path on the reference's own numeric substrate -- TEST / BASELINE
INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs, tests/).

/root/reference is not present on the GPU box, but the third-party wheels that do all of
the reference's arithmetic are (scikit-learn, scipy, numpy).  This module restates the
reference's call sequence on those same library entry points, so that timing it measures
what the reference costs on the box's host cores and its results carry the reference's
ball-tree residue:

    joint MinMaxScaler fit/transform              KDE_distance.py:80-86
    NaN -> 0, KernelDensity(0.03, epanechnikov,   KDE_distance.py:48-56
        ball_tree, euclidean).fit
    exp(score_samples(grid)), grid order          KDE_distance.py:58-64, :17-23
    **gamma, += eps, /= sum                       KDE_distance.py:66-78
    scipy jensenshannon                           KDE_distance.py:129
    optional log(df + eps)                        KDE_distance.py:105-106

`kind` in bench.py's cpu_baseline is "port" for this module.  In the build container
tests/test_oracle.py checks it returns bit-identical values to the unmodified reference.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import jensenshannon
from sklearn.neighbors import KernelDensity
from sklearn.preprocessing import MinMaxScaler


class GridKDEPort:
    """Fixed-bandwidth KDE of one scaled 2-D point set sampled on a uniform R x R grid."""

    def __init__(self, resolution=100, bandwidth=0.03, kernel="epanechnikov", node_stride=1):
        self.resolution = int(resolution)
        self.bandwidth = float(bandwidth)
        self.kernel = kernel
        axis = np.linspace(0, 1, self.resolution)
        # node k = i*R + j sits at (axis[i], axis[j]): column 0 varies slowest
        self.nodes = np.stack(np.meshgrid(axis, axis, indexing="ij"), axis=-1).reshape(-1, 2)
        if node_stride > 1:
            # TIMING ONLY (bench.py --impl reference under a tight time budget): score every
            # node_stride-th node.  The returned number is then not a discrepancy.
            self.nodes = self.nodes[::int(node_stride)]

    def density(self, pts):
        pts = np.where(np.isnan(pts), 0.0, pts)
        est = KernelDensity(bandwidth=self.bandwidth, kernel=self.kernel,
                            algorithm="ball_tree", metric="euclidean").fit(pts)
        return np.exp(est.score_samples(self.nodes))

    def probabilities(self, pts, gamma, eps):
        z = self.density(pts) ** gamma
        z = z + eps
        return z / z.sum()


def js_port(df1, df2, gamma=1.0, eps=1e-12, log=False, resolution=100, bandwidth=0.03,
            kernel="epanechnikov", node_stride=1):
    a = np.asarray(df1, dtype=np.float64)
    b = np.asarray(df2, dtype=np.float64)
    if log:
        a, b = np.log(a + eps), np.log(b + eps)
    scaler = MinMaxScaler().fit(np.concatenate([a, b], axis=0))
    grid = GridKDEPort(resolution, bandwidth, kernel, node_stride)
    p = grid.probabilities(scaler.transform(a), gamma, eps)
    q = grid.probabilities(scaler.transform(b), gamma, eps)
    return np.float64(jensenshannon(p, q))


def _job(args):
    return js_port(*args[0], **args[1])


def js_port_many(pairs, processes=1, **kw):
    """Evaluate independent (df1, df2) pairs; processes>1 uses a fork pool, the reference's
    own scaling mechanism (scripts/pairwise_2D_distance.py:116, distance_to_gridsearch.py:109)."""
    jobs = [(p, kw) for p in pairs]
    if processes <= 1:
        return np.array([_job(j) for j in jobs])
    import multiprocessing as mp

    with mp.get_context("fork").Pool(processes) as pool:
        return np.array(pool.map(_job, jobs))
