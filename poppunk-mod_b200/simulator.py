"""Simulator post-processing around the discrepancy (the caller side of the hot path).

Mirrors, for arrays already in memory or files on disk:

    calculate_gene_freqs   poppunk_mod.py:285-291
    process_result         poppunk_mod.py:351-382   -> [divergence, freq_core, freq_intermediate, freq_rare]
    the ELFI nodes built on it: Distance("euclidean", sim) and Operation(np.log, d)   poppunk_mod.py:571-572

process_results_batch does a whole batch of simulations (BOLFI initial evidence, a Latin-hypercube grid)
with ONE batched discrepancy call; the target is read once instead of once per simulation
(poppunk_mod.py:371 re-reads it for every evaluation).  The gene-frequency summaries are a few thousand
numbers per simulation and stay on the host.
"""
from __future__ import annotations

import os

import numpy as np

from . import KDE_distance as kd
from .sweeps import read_distfile


def calculate_gene_freqs(X, core_threshold, rare_threshold):
    X = np.asarray(X)
    num_genes = (X > 0).sum()
    freq_core = (X >= core_threshold).sum() / num_genes
    freq_rare = np.count_nonzero((0 < X) & (X < rare_threshold)) / num_genes
    freq_intermediate = np.count_nonzero((rare_threshold <= X) & (X < core_threshold)) / num_genes
    return freq_core, freq_intermediate, freq_rare


def _points(x):
    if isinstance(x, (str, os.PathLike)):
        return read_distfile(x).array
    return x


def _freqs(x):
    if isinstance(x, (str, os.PathLike)):
        return np.loadtxt(x, delimiter="\t", dtype="float64")
    return np.asarray(x, dtype=np.float64)


def process_result(simulations, simulations_freqs, obs, gamma, core_threshold, rare_threshold, *, device=None,
                   resolution=100):
    """One simulation: `simulations` / `obs` are (N,2) arrays or TSV paths, `simulations_freqs` a vector or path."""
    fc, fi, fr = calculate_gene_freqs(_freqs(simulations_freqs), core_threshold, rare_threshold)
    div = kd.KDE_JS_divergence(_points(obs), _points(simulations), gamma=gamma, eps=0, log=False, device=device,
                               resolution=resolution)
    return np.array([div, fc, fi, fr])


def process_results_batch(simulations, simulations_freqs, obs, gamma, core_threshold, rare_threshold, *,
                          devices=None, resolution=100):
    """(B,4) summary vectors for B simulations against one target."""
    obs_arr = _points(obs)
    sims = [_points(s) for s in simulations]
    div = kd.KDE_JS_divergence_batch(obs_arr, sims, gamma=gamma, eps=0, log=False, devices=devices,
                                     resolution=resolution)
    out = np.empty((len(sims), 4))
    out[:, 0] = div
    for i, f in enumerate(simulations_freqs):
        out[i, 1:] = calculate_gene_freqs(_freqs(f), core_threshold, rare_threshold)
    return out


def log_euclidean_discrepancy(summaries, observed):
    """log of the euclidean distance between summary vectors and the observed one ([0, fc, fi, fr],
    poppunk_mod.py:494): the quantity BOLFI's GP is fitted to."""
    s = np.atleast_2d(np.asarray(summaries, dtype=np.float64))
    o = np.asarray(observed, dtype=np.float64)
    return np.log(np.sqrt(((s - o) ** 2).sum(axis=1)))
