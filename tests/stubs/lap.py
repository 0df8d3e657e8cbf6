"""Test stand-in for the `lap` package (absent from this image; the reference's trackers import it at module load,
ultralytics/trackers/utils/matching.py:9-17, and would otherwise try to pip-install it).  lapjv with extend_cost / cost_limit
as `lap` defines it -- the cost matrix is extended to (n+m) x (n+m) with cost_limit / 2 off the original block and zeros in
the opposite corner, and assignments that land in the extension count as unmatched -- solved with
scipy.optimize.linear_sum_assignment.  Used identically by the reference arm and by ours in tests/test_gpu_track.py."""
import numpy as np
from scipy.optimize import linear_sum_assignment

__version__ = "0.5.12-stub"


def lapjv(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
    cost = np.asarray(cost, np.float64)
    n_rows, n_cols = cost.shape
    n = n_rows + n_cols
    ext = np.empty((n, n), np.float64)
    ext[:] = cost_limit / 2.0 if cost_limit < np.inf else (cost.max() + 1 if cost.size else 1.0)
    ext[n_rows:, n_cols:] = 0
    ext[:n_rows, :n_cols] = cost
    r, c = linear_sum_assignment(ext)
    x = np.full(n_rows, -1, np.int64)
    y = np.full(n_cols, -1, np.int64)
    for i, j in zip(r, c):
        if i < n_rows and j < n_cols:
            x[i], y[j] = j, i
    opt = float(cost[np.nonzero(x >= 0)[0], x[x >= 0]].sum())
    return (opt, x, y) if return_cost else (x, y)
