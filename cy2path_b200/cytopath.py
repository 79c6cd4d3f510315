"""Mirror of the lineage-clustering step of cy2path/cytopath.py (reference :30-180).

The O(C^2 len^2 d) part — the pairwise dynamic-time-warping or Hausdorff distance matrix between the sampled
Markov chains — and the gather ``cell_state_repr[markov_chains]`` that feeds it (reference :90) run on the
device (csrc/trajdist.cu); the clustering of the resulting C x C matrix stays with scikit-learn / scipy as in the
reference. dtaidistance's ``KMedoids`` is not available here: ``method='kmedoids'`` raises.
"""
from __future__ import annotations

import collections.abc
import logging

import numpy as np

from . import _lib

logger = logging.getLogger("cy2path_b200")


def _device():
    torch = _lib._torch()
    return torch.device("cuda", torch.cuda.current_device())


def _series_on_device(simulations):
    torch = _lib._torch()
    if torch.is_tensor(simulations) and simulations.is_cuda:
        return simulations.to(torch.float64)
    return _lib.to_device(simulations if torch.is_tensor(simulations) else
                          np.ascontiguousarray(simulations, dtype=np.float64), torch.float64, _device())


def compute_dtw_distance_matrix(simulations):
    """``dtaidistance.dtw_ndim.distance_matrix_fast(simulations)`` (reference :92-94): simulations [C, len, d]
    (numpy or cuda tensor) -> float64 numpy [C, C], symmetric, zero diagonal."""
    return _lib.to_host(_lib.trajectory_distance_matrix(_series_on_device(simulations), "dtw"))


def compute_Hausdorff_distance(simulations, distance="euclidean"):
    """reference :30-45 -> float64 numpy [C, C]. Only the Euclidean point distance is implemented."""
    if distance != "euclidean":
        raise NotImplementedError("only distance='euclidean' runs on the device")
    return _lib.to_host(_lib.trajectory_distance_matrix(_series_on_device(simulations), "hausdorff"))


def _cell_state_representation(adata, basis):
    """reference :65-87."""
    from scipy.sparse import issparse

    if isinstance(basis, str):
        return np.asarray(adata.obsm["X_{}".format(basis)])
    X = adata.X.toarray() if issparse(adata.X) else np.asarray(adata.X)
    if basis is None:
        return X
    if isinstance(basis, (collections.abc.Sequence, np.ndarray)):
        genes = np.intersect1d(adata.var.index.values, basis)
        if genes.shape[0] > 0:
            return X[:, adata.var.index.get_indexer(genes)]
    raise ValueError("Cannot compute distances with provided basis key.")


def trajectory_series(adata, basis="pca", differencing=False):
    """The [C, len, d] float64 trajectories of the sampled chains as a cuda tensor: the gather of reference :90
    on the device (plus dtaidistance's first differences when ``differencing``)."""
    torch = _lib._torch()
    chains = np.asarray(adata.uns["markov_chain_sampling"]["state_indices"])
    rep = _lib.to_device(np.ascontiguousarray(_cell_state_representation(adata, basis), dtype=np.float64),
                         torch.float64, _device())
    idx = torch.as_tensor(chains.astype(np.int64)).to(rep.device)
    series = rep[idx]  # [C, len, d]
    if differencing:
        series = series[:, 1:] - series[:, :-1]
    return series.contiguous()


def cluster_markov_chains(adata, num_lineages=None, method="HDBSCAN", distance_func="dtw", differencing=False,
                          basis="pca", max_lineages=10, n_jobs=-1):
    """reference :48-180 -> (cluster_labels, model). ``method``: 'HDBSCAN' (default) or 'linkage' (Ward linkage on
    the condensed distances, as dtaidistance's LinkageTree builds it; model.linkage holds the scipy linkage)."""
    from scipy.cluster import hierarchy
    from scipy.spatial.distance import squareform

    num_chains = adata.uns["markov_chain_sampling"]["sampling_params"]["num_chains"]
    logger.info("Clustering the samples.")
    series = trajectory_series(adata, basis=basis, differencing=differencing)
    if callable(distance_func):
        distances = np.asarray(distance_func(series.cpu().numpy()))
    elif distance_func in ("dtw", "hausdorff"):
        distances = _lib.to_host(_lib.trajectory_distance_matrix(series, distance_func))
    else:
        raise ValueError("distance_func must be 'dtw', 'hausdorff' or a callable")
    if "cytopath" not in adata.uns:
        adata.uns["cytopath"] = {}
    adata.uns["cytopath"]["trajectory_distances"] = distances

    class _Linkage:
        def __init__(self, Z):
            self.linkage = Z

    def cluster_n(k):
        if k <= 0:
            raise ValueError("num_lineages must be greater than 0!")
        if method == "linkage":
            Z = hierarchy.linkage(squareform(distances, checks=False), method="ward")
            return hierarchy.fcluster(Z, k, criterion="maxclust"), _Linkage(Z)
        if method == "kmedoids":
            raise NotImplementedError("method='kmedoids' needs dtaidistance.clustering.KMedoids (not available)")
        raise ValueError("Invalid method for clustering!")

    if method == "HDBSCAN":
        from sklearn.cluster import HDBSCAN

        if num_lineages is not None:
            logger.warning("num_lineages ignored for method HDBSCAN!")
        model = HDBSCAN(min_cluster_size=max(2, int(num_chains * 0.05)), metric="precomputed", n_jobs=n_jobs,
                        allow_single_cluster=True)
        return model.fit_predict(distances), model
    if type(num_lineages) is int and method in ("linkage", "kmedoids"):
        return cluster_n(num_lineages)
    if num_lineages == "auto" and method in ("linkage", "kmedoids"):
        from sklearn.metrics import silhouette_score

        if max_lineages <= 2:
            raise ValueError("max_lineages must be greater than 2 for auto mode with linkage or kmedoids!")
        scores, labels, models = {}, {}, {}
        for k in range(2, max_lineages + 1):
            labels[k], models[k] = cluster_n(k)
            scores[k] = silhouette_score(distances, labels[k], metric="precomputed")
        best = max(scores, key=scores.get)
        logger.info("Optimal number of lineages is %d with silhouette score %.3g", best, scores[best])
        adata.uns["cytopath"]["silhouette_scores"] = scores
        adata.uns["cytopath"]["optimal_num_lineages"] = best
        return labels[best], models[best]
    raise ValueError("Incompatible num_lineages and method specification!")
