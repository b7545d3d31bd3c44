"""Read-back helpers over the library's scratch buffers (layout: csrc/lgs_common.cuh).  Used by bench.py to
count how many list entries a frame really sorted / replayed (the lazy sort makes that data-dependent) and
by the full-size tests for white-box properties.  Everything stays on the device (torch views)."""
import torch

NB = 64  # LGS_NB, depth buckets per bin


def _al(x):
    return (x + 255) & ~255


def effective_rows_per_bin(H, rows_per_bin=0):
    rb = rows_per_bin if rows_per_bin in (1, 2, 4, 8, 16) else 8
    while rb > 1 and rb > H:
        rb >>= 1
    return rb


def _u32(t):
    """uint8 slice -> int64 tensor holding the unsigned 32-bit values."""
    return t.view(torch.int32).long() & 0xffffffff


def frame_views(frame, P, H, W, rows_per_bin=0):
    """Decode the three scratch buffers of a capi.Frame after forward()."""
    RB = effective_rows_per_bin(H, rows_per_bin)
    gx = (W + 15) // 16
    nrg = (H + RB - 1) // RB
    nbins = gx * nrg
    gb, img, bb = frame.geom, frame.image, frame.binning
    o = 0
    rec = gb[o:o + 64 * P].view(torch.float32).view(P, 16); o = _al(o + 64 * P)
    aux = _u32(gb[o:o + 16 * P]).view(P, 4); o = _al(o + 16 * P)
    o = _al(o + nbins * NB * 4)  # cnt
    loc = _u32(gb[o:o + nbins * NB * 4]).view(nbins, NB); o = _al(o + nbins * NB * 4)
    binbase = _u32(gb[o:o + (nbins + 1) * 4]); o = _al(o + (nbins + 1) * 4)
    o = 0
    final_T = img[o:o + 4 * H * W].view(torch.float32).view(H, W); o = _al(o + 4 * H * W)
    n_contrib = _u32(img[o:o + 4 * H * W]).view(H, W); o = _al(o + 4 * H * W)
    sorted_end = _u32(img[o:o + 4 * nbins])
    N = int(binbase[-1].item())
    entries = _u32(bb[:16 * N]).view(N, 4) if N else torch.zeros((0, 4), dtype=torch.long, device=gb.device)
    # deepest contributor per bin = what the backward pass replays
    Hp, Wp = nrg * RB, gx * 16
    pad = torch.zeros((Hp, Wp), dtype=torch.long, device=img.device)
    pad[:H, :W] = n_contrib
    per_bin = pad.view(nrg, RB, gx, 16).permute(0, 2, 1, 3).reshape(nbins, -1).max(dim=1).values
    consumed = dict(nbins=nbins, rows_per_bin=RB, sorted=int(sorted_end.sum().item()), replayed=int(per_bin.sum().item()),
                    sorted_max_bin=int(sorted_end.max().item()), replayed_max_bin=int(per_bin.max().item()),
                    blended_pairs_upper=int(n_contrib.sum().item()))
    return dict(rec=rec, aux=aux, loc=loc, binbase=binbase, final_T=final_T, n_contrib=n_contrib, sorted_end=sorted_end,
                entries=entries, consumed=consumed, nbins=nbins, rows_per_bin=RB)


def consumed_entries(frame, H, W, rows_per_bin=0, P=None):
    P = frame.out["radii"].shape[0] if P is None else P
    return frame_views(frame, P, H, W, rows_per_bin)["consumed"]
