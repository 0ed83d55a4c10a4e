"""Oracle (test infrastructure): the tensor arithmetic of the reference's evaluation loops that
SURVEY.md 8 row f2 moves onto the device.  CPU only; never imported by the product path.

  * spn_top_classes  -- src/core/inference.py:180-181: torch.topk(weights, num_neighbors, dim=1) followed
    by torch.softmax over the k winners (values sorted descending; int64 indices).
  * krn_keypoints_pix -- src/core/inference.py:236-243 (_keypts_to_pose before pnp): keypoints in the
    crop's [0,1] frame -> pixels, `x * (xmax - xmin) + xmin` evaluated by numpy in float32.

Pinned by tests/golden/postproc.npz, captured from the unmodified reference loops
(oracle/make_golden_postproc.py).
"""
import numpy as np


def spn_top_classes(weights, k):
    """weights: [B,N] float32 array -> (topWeights [B,k] float32, topClasses [B,k] int64)."""
    w = np.asarray(weights, dtype=np.float32)
    # stable descending order: ties resolve to the lowest index
    order = np.argsort(-w, axis=1, kind='stable')[:, :k]
    top = np.take_along_axis(w, order, axis=1)
    e = np.exp((top - top[:, :1]).astype(np.float32))
    return (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32), order.astype(np.int64)


def krn_keypoints_pix(x_pr, y_pr, bbox):
    """x_pr, y_pr: [B,K] float32; bbox: [B,4] float32 (xmin, xmax, ymin, ymax) -> [B,K,2] float32 pixels."""
    x, y, bb = (np.asarray(a, dtype=np.float32) for a in (x_pr, y_pr, bbox))
    out = np.empty(x.shape + (2,), np.float32)
    for b in range(x.shape[0]):
        xmin, xmax, ymin, ymax = bb[b]
        out[b, :, 0] = x[b] * (xmax - xmin) + xmin
        out[b, :, 1] = y[b] * (ymax - ymin) + ymin
    return out
