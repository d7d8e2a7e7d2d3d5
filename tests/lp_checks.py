"""Size-independent checks shared by the CPU and GPU tests."""
import numpy as np


def certify_optimal(A, b, c, rhs, obj_row, basis, rel=1e-8):
    """Size-independent proof that a solved tableau is THE optimum of max c.x, Ax <= b, x >= 0:
    primal feasibility, dual feasibility and a closed duality gap.  The optimum value is unique,
    so this pins the objective to within `rel` of the reference's exact-rational answer without
    running the reference (BASELINE north_star: objective within 1e-8)."""
    m, n = A.shape
    x = np.zeros(n + m)
    x[basis] = rhs[:m]
    xs, slack = x[:n], x[n:]
    scale = max(1.0, float(np.abs(b).max()))
    assert (x >= -1e-9 * scale).all()                               # x, slack >= 0
    assert np.abs(A @ xs + slack - b).max() <= 1e-9 * scale * n     # Ax + s = b
    y = obj_row[n:n + m]                                            # duals = slack reduced costs
    d = obj_row[:n]
    assert (y >= -1e-9).all() and (d >= -1e-9).all()                # dual feasible
    assert np.abs(A.T @ y - c - d).max() <= 1e-9 * max(1.0, float(np.abs(c).max())) * m
    primal, dual, reported = float(c @ xs), float(b @ y), float(rhs[m])
    assert abs(primal - reported) <= rel * abs(reported)
    assert abs(dual - reported) <= rel * abs(reported)              # strong duality: optimal
    return reported
