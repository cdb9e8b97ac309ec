"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's fold-in math:

  Generation.recomputeSolver                    online/src/net/myrrix/online/generation/Generation.java:141-158
  ServerRecommender.updateFeatures              online/src/net/myrrix/online/ServerRecommender.java:865-907
  ServerRecommender.foldInWeight                :981-994
  ServerRecommender.buildAnonymousUserFeatures  :561-608 (arithmetic only)

The k x k solves use numpy LU in fp64 -- NOT the RRQR
restatement the product uses: an independent route to the same well-conditioned answer.
Parity unpinned: the reference has no unit test with inline data for these methods.
"""
import numpy as np


class IllConditioned(Exception):
    pass


def fold_in_weight(estimate, value, learn_rate=1.0):
    assert np.isfinite(estimate)
    if value > 0.0 and estimate < 1.0:
        w = (1.0 - 1.0 / (1.0 + float(np.float32(value)))) * (1.0 - max(0.0, estimate))
    elif value < 0.0 and estimate > 0.0:
        w = (1.0 - 1.0 / (1.0 - float(np.float32(value)))) * -min(1.0, estimate)
    else:
        w = 0.0
    return learn_rate * w


def dot_f(x, y):
    """SimpleVectorMath.dot: float products (rounded to fp32) summed in fp64."""
    return float(np.sum((x.astype(np.float32) * y.astype(np.float32)).astype(np.float64)))


def check_mtm(MTM):
    inf_norm = np.abs(MTM).sum(axis=1).max()   # RealMatrix.getNorm(): maximum absolute row sum
    if inf_norm < 1.0:
        raise IllConditioned(inf_norm)


def update_features(user, item, value, xtx, yty, learn_rate=1.0):
    """Returns the new (user, item) rows (float32); either Gramian may be None."""
    w = fold_in_weight(dot_f(user, item), value, learn_rate)
    user, item = user.copy(), item.copy()
    if w == 0.0:
        return user, item
    item_fold = None if xtx is None else np.linalg.solve(xtx, user.astype(np.float64))
    user_fold = None if yty is None else np.linalg.solve(yty, item.astype(np.float64))
    if item_fold is not None:
        item = (item + (w * item_fold).astype(np.float32)).astype(np.float32)
    if user_fold is not None:
        user = (user + (w * user_fold).astype(np.float32)).astype(np.float32)
    return user, item


def anonymous_user(item_rows, values, yty, learn_rate=1.0):
    out = np.zeros(item_rows.shape[1], np.float32)
    for j, row in enumerate(item_rows):
        fold = np.linalg.solve(yty, row.astype(np.float64))
        w = fold_in_weight(0.0, 1.0 if values is None else values[j], learn_rate)
        if w != 0.0:
            out = (out + (w * fold).astype(np.float32)).astype(np.float32)
    return out
