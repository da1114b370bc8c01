"""Shared helpers of the parity tests."""
import numpy as np

# Force parity bar (BASELINE.json north_star): 1e-5 relative, fp32.  "Relative" is taken per
# atom against the sum of the magnitudes of the pair forces acting on it (hard parts):
# the net force itself can cancel to ~0, which no fp32 summation can resolve to 1e-5.
FORCE_RTOL = 1e-5
ENERGY_RTOL = 1e-5


def force_rel_err(f_test, f_truth, sumabs):
    scale = np.maximum(sumabs, 1e-3 * max(float(sumabs.max()), 1e-30))
    return np.abs(f_test[:, :3].astype(np.float64) - f_truth[:, :3].astype(np.float64)).max(1) / scale


def csr_rows(start, idx):
    return [idx[start[i]:start[i + 1]] for i in range(len(start) - 1)]


def numpy_row(xyzq, i, ext, periodic, r_list):
    """Neighbour row of atom i by brute force in numpy fp32, same expression as the oracle
    (oracle/md_oracle.c dist2_f32) -- used at sizes where the O(N^2) oracle is too slow."""
    x = xyzq[:, :3].astype(np.float32)
    d = x[i][None, :] - x
    if periodic:
        e = np.asarray(ext, np.float32)[None, :]
        d = d - np.rint(d / e) * e
    r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    rl = np.float32(r_list)
    hit = r2 < rl * rl
    hit[i] = False
    return np.nonzero(hit)[0].astype(np.int32)
