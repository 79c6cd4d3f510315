"""TEST INFRASTRUCTURE — CPU restatement of the trajectory distance matrices of the cytopath workflow
(reference cy2path/cytopath.py:30-45, 90-96, 122).

Not product code (see oracle/msm.py header for who may import this).

PARITY UNPINNED for this file: the arithmetic lives in two third-party packages the reference lists without a
version pin (pyproject.toml:39-40, ``hausdorff`` and ``dtaidistance``) and neither is installed in this image
(no network), so their outputs cannot be generated here. What is restated is their published algorithm:

* ``dtaidistance.dtw_ndim.distance_matrix_fast(series)`` with default settings (no window, no pruning, no
  penalty): classic dynamic time warping between two equal-dimension series with the SQUARED Euclidean distance
  between points as local cost, D[i,j] = cost(i,j) + min(D[i-1,j], D[i,j-1], D[i-1,j-1]), distance =
  sqrt(D[len_a, len_b]); the matrix is symmetric with a zero diagonal.
* ``hausdorff.hausdorff_distance(A, B, distance='euclidean')``: max( max_i min_j |a_i - b_j|,
  max_j min_i |a_i - b_j| ), filled symmetrically by cytopath.py:30-45.
* ``dtaidistance.preprocessing.differencing(series)``: first differences along the time axis.

Cross-checks that ARE possible here (tests/test_oracle.py): the Hausdorff restatement equals scipy's own
``scipy.spatial.distance.directed_hausdorff`` (max of the two directions) to 1e-12, and the DTW recurrence equals the
cheapest monotone lattice path found by scipy's Dijkstra on the explicit DTW graph. The two packages' own outputs
remain unchecked, hence "unpinned".
"""
import numpy as np


def dtw_distance(a, b):
    """a [la, d], b [lb, d] float64 -> DTW distance (sqrt of the accumulated squared Euclidean cost)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    la, lb = a.shape[0], b.shape[0]
    cost = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    D = np.full((la + 1, lb + 1), np.inf)
    D[0, 0] = 0.0
    for i in range(1, la + 1):
        row, prev = D[i], D[i - 1]
        for j in range(1, lb + 1):
            row[j] = cost[i - 1, j - 1] + min(prev[j], row[j - 1], prev[j - 1])
    return float(np.sqrt(D[la, lb]))


def dtw_matrix(series):
    """series [C, len, d] -> symmetric [C, C] float64, zero diagonal."""
    series = np.asarray(series, dtype=np.float64)
    C = series.shape[0]
    out = np.zeros((C, C))
    for i in range(C):
        for j in range(i + 1, C):
            out[i, j] = out[j, i] = dtw_distance(series[i], series[j])
    return out


def hausdorff_distance(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    dist = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
    return float(max(dist.min(1).max(), dist.min(0).max()))


def hausdorff_matrix(series):
    """reference cytopath.py:30-45 (compute_Hausdorff_distance)."""
    series = np.asarray(series, dtype=np.float64)
    C = series.shape[0]
    out = np.zeros((C, C))
    for i in range(C - 1, -1, -1):
        for j in range(i):
            out[i, j] = out[j, i] = hausdorff_distance(series[i], series[j])
    return out


def differencing(series):
    return np.diff(np.asarray(series, dtype=np.float64), axis=1)
