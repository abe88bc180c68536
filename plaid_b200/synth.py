"""Deterministic synthetic inputs of the shapes BASELINE.json names.

Two back-ends with the same sampling scheme: numpy (host; tests, CPU baseline) and torch
(device; the 125k-cell benchmark shard is generated directly in HBM).  Seeds are
20261017 + config index.  The streams of the two back-ends differ (PCG64 vs Philox); every
consumer compares GPU and oracle on the SAME generated arrays, never across back-ends.

  gene sets  "MSigDB-scale": size_s = clip(round(exp(N(ln 60, 0.9))), 5, 2000), members drawn
             without replacement with gene weight proportional to rank^-0.5.
  sparse X   pbmc3k-shaped: nnz_j = clip(round(exp(N(ln(0.07 P), 0.35))), 0.015 P, 0.25 P), genes
             drawn without replacement with weight proportional to rank^-0.9, counts
             k = 1 + Geometric(0.55), value = ln(1 + k * 1e4 / sum_k)  (Seurat LogNormalize ->
             heavy exact ties).
  dense X    x_gj = mu_g + sigma_g * N(0,1), mu_g ~ U(2,14), sigma_g ~ U(0.3,1.5).

Weighted sampling without replacement uses the Gumbel-top-k identity (perturb log-weights
with Gumbel noise, keep the k largest), batched so that no intermediate exceeds ~1 GB.
"""
from __future__ import annotations

import numpy as np

SEED0 = 20261017


def gene_names(P: int):
    return [f"G{k:05d}" for k in range(P)]


def set_names(S: int):
    return [f"SET{k:05d}" for k in range(S)]


# --------------------------------------------------------------------------------------
# numpy back-end
# --------------------------------------------------------------------------------------
def _topk_mask_np(rng, logw, k, batch=256):
    """rows of a boolean (len(k), P) mask with exactly k[r] True, sampled w/o replacement."""
    P = logw.size
    n = k.size
    rows, cols = [], []
    for b0 in range(0, n, batch):
        kb = k[b0:b0 + batch]
        g = rng.gumbel(size=(kb.size, P)) + logw[None, :]
        order = np.argsort(-g, axis=1, kind="stable")
        for r in range(kb.size):
            sel = np.sort(order[r, :kb[r]])
            rows.append(np.full(sel.size, b0 + r, dtype=np.int64))
            cols.append(sel)
    return np.concatenate(rows) if rows else np.zeros(0, np.int64), \
        np.concatenate(cols) if cols else np.zeros(0, np.int64)


def genesets_numpy(P: int, S: int, seed: int = SEED0, size_cap=(5, 2000)):
    """csc_matrix P x S of ones."""
    import scipy.sparse as sp
    rng = np.random.Generator(np.random.PCG64(seed))
    sizes = np.clip(np.rint(np.exp(rng.normal(np.log(60.0), 0.9, size=S))), size_cap[0],
                    min(size_cap[1], P)).astype(np.int64)
    logw = -0.5 * np.log(np.arange(1, P + 1))
    sets, genes = _topk_mask_np(rng, logw, sizes)
    G = sp.csc_matrix((np.ones(genes.size), (genes, sets)), shape=(P, S))
    G.sort_indices()
    return G


def sparse_x_numpy(P: int, N: int, seed: int = SEED0, density: float = 0.07):
    """csc_matrix P x N, pbmc3k-shaped."""
    import scipy.sparse as sp
    rng = np.random.Generator(np.random.PCG64(seed))
    nnz = np.clip(np.rint(np.exp(rng.normal(np.log(density * P), 0.35, size=N))), max(1, int(0.015 * P)),
                  max(1, int(0.25 * P))).astype(np.int64)
    logw = -0.9 * np.log(np.arange(1, P + 1))
    cells, genes = _topk_mask_np(rng, logw, nnz)
    k = rng.geometric(0.55, size=genes.size).astype(np.float64)  # support 1, 2, ...
    tot = np.bincount(cells, weights=k, minlength=N)
    val = np.log1p(k * 1e4 / tot[cells])
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(nnz, out=indptr[1:])
    X = sp.csc_matrix((val, genes.astype(np.int32), indptr), shape=(P, N))
    X.sort_indices()
    return X


def dense_x_numpy(P: int, N: int, seed: int = SEED0):
    rng = np.random.Generator(np.random.PCG64(seed))
    mu = rng.uniform(2, 14, size=P)
    sg = rng.uniform(0.3, 1.5, size=P)
    return np.asfortranarray(mu[:, None] + sg[:, None] * rng.standard_normal((P, N)))


# --------------------------------------------------------------------------------------
# torch back-end (device-resident generation for the large benchmark shard)
# --------------------------------------------------------------------------------------
def _gumbel_select_torch(gen, logw, k, batch):
    """Yield (row_ids, col_ids) of exactly k[r] picks per row, columns ascending per row."""
    import torch
    P = logw.numel()
    n = k.numel()
    dev = logw.device
    ar = torch.arange(P, device=dev)
    for b0 in range(0, n, batch):
        kb = k[b0:b0 + batch]
        u = torch.rand((kb.numel(), P), generator=gen, device=dev, dtype=torch.float32)
        g = logw[None, :] - torch.log(-torch.log(u.clamp_(1e-20, 1.0 - 1e-7)))
        order = torch.argsort(g, dim=1, descending=True)
        del g, u
        mask = torch.zeros((kb.numel(), P), dtype=torch.bool, device=dev)
        mask.scatter_(1, order, ar[None, :] < kb[:, None])
        del order
        rc = mask.nonzero(as_tuple=False)  # row-major: rows ascending, cols ascending
        yield rc[:, 0] + b0, rc[:, 1]


def genesets_torch(P: int, S: int, seed: int = SEED0, device="cuda", batch=1024):
    """Returns host numpy (Gp int32[S+1], Gi int32[nnz]) generated on `device`."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    sizes = torch.clamp(torch.round(torch.exp(torch.randn(S, generator=gen, device=device) * 0.9 + np.log(60.0))),
                        5, min(2000, P)).to(torch.int64)
    logw = -0.5 * torch.log(torch.arange(1, P + 1, device=device, dtype=torch.float32))
    gi = [c.to(torch.int32) for _, c in _gumbel_select_torch(gen, logw, sizes, batch)]
    Gi = torch.cat(gi).cpu().numpy()
    Gp = np.zeros(S + 1, dtype=np.int32)
    np.cumsum(sizes.cpu().numpy(), out=Gp[1:])
    return Gp, Gi


def sparse_x_torch(P: int, N: int, seed: int = SEED0, device="cuda", density: float = 0.07, batch=2048):
    """Device CSC (p int32[N+1], i int32[nnz], x float64[nnz]) generated in HBM."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    nnz = torch.clamp(torch.round(torch.exp(torch.randn(N, generator=gen, device=device) * 0.35 + np.log(density * P))),
                      max(1, int(0.015 * P)), max(1, int(0.25 * P))).to(torch.int64)
    logw = -0.9 * torch.log(torch.arange(1, P + 1, device=device, dtype=torch.float32))
    tot_nnz = int(nnz.sum().item())
    xi = torch.empty(tot_nnz, dtype=torch.int32, device=device)
    xx = torch.empty(tot_nnz, dtype=torch.float64, device=device)
    p = torch.zeros(N + 1, dtype=torch.int64, device=device)
    p[1:] = torch.cumsum(nnz, 0)
    off = 0
    for cells, genes in _gumbel_select_torch(gen, logw, nnz, batch):
        m = genes.numel()
        u = torch.rand(m, generator=gen, device=device, dtype=torch.float64)
        k = torch.floor(torch.log(1.0 - u) / np.log(1.0 - 0.55)) + 1.0  # Geometric(0.55) on 1, 2, ...
        tot = torch.zeros(N, dtype=torch.float64, device=device).index_add_(0, cells, k)
        xi[off:off + m] = genes.to(torch.int32)
        xx[off:off + m] = torch.log1p(k * 1e4 / tot[cells])
        off += m
    assert off == tot_nnz
    return p.to(torch.int32), xi, xx
