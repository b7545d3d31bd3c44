"""TEST INFRASTRUCTURE -- numpy restatement of GaussianModel.training_statis (scene/gaussian_model.py:597-618).
PARITY PIN: tests/golden/gt*.npz (the reference's own method, exec()'d unmodified by oracle/make_goldens_statis.py)."""
import numpy as np


def training_statis(acc, K, grad, opacity, update_filter, selection_mask, anchor_visible):
    """acc: dict of the four accumulators (copied, not modified).  Returns the updated copies."""
    out = {k: v.copy() for k, v in acc.items()}
    op = np.maximum(opacity.reshape(-1), 0).reshape(-1, K)                                 # :599-603
    out["opacity_accum"][anchor_visible] += op.sum(1, keepdims=True)                       # :604
    out["anchor_demon"][anchor_visible] += 1                                               # :606
    vis_k = np.repeat(anchor_visible, K)                                                   # :609
    combined = np.zeros(vis_k.shape[0], bool)
    combined[vis_k] = selection_mask.reshape(-1)                                           # :610-611
    tmp = combined.copy()
    combined[tmp] = update_filter.reshape(-1)                                              # :612-613
    norm = np.sqrt((grad[update_filter.reshape(-1), 2:] ** 2).sum(1, keepdims=True))       # :616
    out["offset_gradient_accum"][combined] += norm                                         # :617
    out["offset_denom"][combined] += 1                                                     # :618
    return out
