"""Seeded synthetic inputs generated ON the device (torch), for the configurations whose host-side
generation (synth.py, NumPy) would take minutes and tens of GB per rank: 10M x 128 (config 3) and
the 12.5M-node per-GPU slice of the 100M-node graph (config 5).

Same graph model as synth.hnsw_like_graph: level-0 out-degree U{M..2M}, level l >= 1
U{M/2..M} over the ~n/M^l nodes present, a fraction ``locality`` of the neighbours from a +-window id
range and the rest uniform over all nodes, 1% isolated nodes, no self loops, no duplicate neighbour in
a row, ascending U(0,1) edge weights per row.  The streams differ from the NumPy generator's (torch
Philox on the device), so graphs are reproducible per (seed, shape, device type), not equal to synth's.

Nothing here computes the layer; torch is plumbing (device memory + RNG).
"""
from __future__ import annotations

from typing import Optional, Tuple


def _gen(seed: int, device):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def make_features(n: int, d: int, seed: int = 0x5EED, device="cuda", row_offset: int = 0, slab: int = 1 << 21):
    """N(0,1) rows, L2-normalised.  Row r of the block depends on (seed, slab index of row_offset + r)
    only through the slab it lives in: blocks are generated slab by slab from seed + slab id, so any
    rank can regenerate any aligned range of another rank's rows."""
    import torch

    if row_offset % slab:
        raise ValueError("row_offset must be a multiple of the slab size")
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    for a in range(0, n, slab):
        b = min(a + slab, n)
        g = _gen(seed * 1000003 + (row_offset + a) // slab, device)
        x = torch.randn((slab, d), generator=g, dtype=torch.float32, device=device)  # whole slabs: see feature_rows
        x /= x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        out[a:b] = x[: b - a]
    return out


def hnsw_like_graph(n: int, m: int, seed: int = 0xBEEF, locality: float = 0.0, window: int = 4096,
                    iso_frac: float = 0.01, level: int = 0, n_total: Optional[int] = None, row_offset: int = 0,
                    device="cuda", slab: int = 1 << 21) -> Tuple["torch.Tensor", "torch.Tensor", "torch.Tensor"]:
    """Rows ``[row_offset, row_offset + n)`` of an ``n_total``-node graph with GLOBAL column ids.
    Returns device tensors ``(row_ptr int64 [n+1], col_idx int32 [E], edge_w float32 [E])``."""
    import torch

    nt = n if n_total is None else n_total
    if nt >= 2 ** 31:
        raise ValueError("node ids must fit int32")
    lo, hi = (m, 2 * m) if level == 0 else (max(m // 2, 1), m)
    hi = min(hi, max(nt - 1, 0))
    lo = min(lo, hi)
    width = hi
    cols_out, w_out, deg_out = [], [], []
    for a in range(0, n, slab):
        b = min(a + slab, n)
        r = b - a
        g = _gen(seed * 1000003 + level * 7919 + (row_offset + a) // slab, device)
        deg = torch.randint(lo, hi + 1, (r,), generator=g, device=device, dtype=torch.int32) if hi > 0 else \
            torch.zeros(r, dtype=torch.int32, device=device)
        if level > 0:
            present = torch.rand(r, generator=g, device=device) < float(m) ** (-level)
            deg = torch.where(present, deg, torch.zeros_like(deg))
        if iso_frac > 0:
            deg = torch.where(torch.rand(r, generator=g, device=device) < iso_frac, torch.zeros_like(deg), deg)
        if width == 0 or int(deg.max()) == 0:
            deg_out.append(deg.to(torch.int64))
            continue
        rowid = (row_offset + a + torch.arange(r, device=device, dtype=torch.int32))[:, None]
        # only the rows that have neighbours are drawn (upper levels: ~n / M^level of them)
        live = torch.nonzero(deg > 0).squeeze(1)
        rid = rowid[live]
        dlive = deg[live]

        def draw(rows_id):
            k = rows_id.shape[0]
            far = torch.randint(0, nt - 1, (k, width), generator=g, device=device, dtype=torch.int32)
            far += (far >= rows_id).to(torch.int32)  # skip self
            if locality <= 0:
                return far
            w = min(window, max((nt - 1) // 2, 1))
            off = torch.randint(1, w + 1, (k, width), generator=g, device=device, dtype=torch.int32)
            sign = torch.randint(0, 2, (k, width), generator=g, device=device, dtype=torch.int32) * 2 - 1
            near = torch.remainder(rows_id + off * sign, nt)
            near = torch.where(near == rows_id, torch.remainder(near + 1, nt), near)
            pick = torch.rand((k, width), generator=g, device=device) < locality
            return torch.where(pick, near, far)

        cols = draw(rid)
        pos = torch.arange(width, device=device, dtype=torch.int32)[None, :]
        mask = pos < dlive[:, None]
        for _ in range(64):  # re-draw the rows that hold a duplicate among their first deg entries
            srt = torch.sort(torch.where(mask, cols, -1 - pos), dim=1).values
            bad = torch.nonzero((srt[:, 1:] == srt[:, :-1]).any(dim=1)).squeeze(1)
            if bad.numel() == 0:
                break
            cols[bad] = draw(rid[bad])
        else:
            raise RuntimeError("could not draw duplicate-free neighbour rows")
        wts = torch.sort(torch.rand((live.numel(), width), generator=g, device=device), dim=1).values
        cols_out.append(cols[mask])
        w_out.append(wts[mask])
        deg_out.append(deg.to(torch.int64))
        del cols, wts, mask, srt
    deg_all = torch.cat(deg_out) if deg_out else torch.zeros(0, dtype=torch.int64, device=device)
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg_all, 0, out=row_ptr[1:])
    col = torch.cat(cols_out) if cols_out else torch.zeros(0, dtype=torch.int32, device=device)
    w = torch.cat(w_out) if w_out else torch.zeros(0, dtype=torch.float32, device=device)
    return row_ptr, col.contiguous(), w.contiguous()


def feature_rows(rows, d: int, seed: int = 0x5EED, device="cuda", slab: int = 1 << 21):
    """The rows ``rows`` (indices in the slab-addressed row space of :func:`make_features`, i.e.
    ``row_offset + r``) regenerated on their own: only the slabs that hold them are drawn."""
    import numpy as np
    import torch

    rows = np.asarray(rows, dtype=np.int64)
    out = torch.empty((rows.size, d), dtype=torch.float32, device=device)
    for sid in np.unique(rows // slab):
        sel = np.nonzero(rows // slab == sid)[0]
        g = _gen(seed * 1000003 + int(sid), device)
        x = torch.randn((slab, d), generator=g, dtype=torch.float32, device=device)
        x /= x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        out[torch.as_tensor(sel, device=device)] = x[torch.as_tensor(rows[sel] - sid * slab, device=device)]
        del x
    return out
