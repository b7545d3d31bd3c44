"""Read-back helpers over the library's scratch buffers (layout: csrc/lgs_common.cuh).  Used by bench.py to
count how many list entries a frame really sorted / replayed (the lazy sort makes that data-dependent)."""
import torch


def _al(x):
    return (x + 255) & ~255


def effective_rows_per_bin(H, rows_per_bin=0):
    rb = rows_per_bin if rows_per_bin in (1, 2, 4, 8, 16) else 8
    while rb > 1 and rb > H:
        rb >>= 1
    return rb


def consumed_entries(frame, H, W, rows_per_bin=0):
    RB = effective_rows_per_bin(H, rows_per_bin)
    gx = (W + 15) // 16
    nrg = (H + RB - 1) // RB
    nbins = gx * nrg
    img = frame.image
    o = _al(4 * H * W)
    n_contrib = img[o:o + 4 * H * W].view(torch.int32).view(H, W)
    o = _al(o + 4 * H * W)
    sorted_end = img[o:o + 4 * nbins].view(torch.int32)
    # deepest contributor per bin = what the backward pass replays
    Hp, Wp = nrg * RB, gx * 16
    pad = torch.zeros((Hp, Wp), dtype=torch.int32, device=img.device)
    pad[:H, :W] = n_contrib
    per_bin = pad.view(nrg, RB, gx, 16).permute(0, 2, 1, 3).reshape(nbins, -1).max(dim=1).values
    return dict(nbins=nbins, rows_per_bin=RB, sorted=int(sorted_end.sum().item()), replayed=int(per_bin.sum().item()),
                blended_pairs_upper=int(n_contrib.sum().item()))
