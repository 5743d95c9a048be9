"""Host-side logic for sharding the cells (columns) of one count matrix over several GPUs.

The VB update shards naturally over cells: given the W-side panel every cell column is independent,
and the only exchange per iteration is the sum over shards of the W-side statistics Sw (genes x r),
the r-vector rowSums(eh) and a handful of scalars.  This module holds the pieces
that do not need a GPU: the nnz-balanced partition and the packing of that exchange buffer, so
that they can be tested with torch.distributed/gloo on CPU.
"""
import numpy as np


def balanced_bounds(colptr, nranks, align=1):
    """Contiguous column ranges with (nearly) equal numbers of nonzeros.

    colptr: CSC column pointers (length m+1).  Returns nranks+1 boundaries; boundary b is the
    column whose start offset is closest to b*nnz/nranks, rounded to a multiple of `align`.
    """
    colptr = np.asarray(colptr, dtype=np.int64)
    m = len(colptr) - 1
    nnz = int(colptr[-1])
    bounds = [0]
    for b in range(1, nranks):
        target = nnz * b / nranks
        j = int(np.searchsorted(colptr, target, side="left"))
        if j > 0 and abs(colptr[j - 1] - target) <= abs(colptr[min(j, m)] - target):
            j -= 1
        if align > 1:
            j = int(round(j / align)) * align
        j = min(max(j, bounds[-1]), m)
        bounds.append(j)
    bounds.append(m)
    return bounds


def shard_csc(csc, c0, c1):
    """Columns [c0, c1) of a scipy CSC matrix as a CSC matrix (all gene rows kept)."""
    return csc[:, c0:c1].tocsc()


def exchange_len(n, r):
    """Doubles in the per-iteration all-reduce buffer: Sw (n*r) + rowSums(eh) (r) + 8 scalars
    (H prior/entropy sum, sum log lh, sum eh, entropy-collapse term of H, sum x log p,
    entropy-collapse term of W, 2 spare)."""
    return n * r + r + 8


def pack_exchange(Sw, ehsum, scalars):
    """Sw: n x r partial statistics of this shard; ehsum: r; scalars: up to 8 partial sums."""
    s = np.zeros(8)
    s[:len(scalars)] = scalars
    return np.concatenate([np.asarray(Sw, dtype=np.float64).ravel(), np.asarray(ehsum, float), s])


def unpack_exchange(buf, n, r):
    buf = np.asarray(buf)
    return buf[:n * r].reshape(n, r), buf[n * r:n * r + r], buf[n * r + r:]
