"""Row-range sharding of the interaction matrix over ranks (one process per GPU).

Users and items are each cut into `world` equal contiguous blocks of ceil(n / world) rows
(the factor replicas are padded to world * block rows so the per-half exchange is one
in-place all-gather of equal blocks, csrc/als_abi.cu: exchange()).  Rank r updates user
block r in the X half and item block r in the Y half; both orientations of R are therefore
sharded by their own row index, exactly the two maps the reference constructor is handed
(RbyRow / RbyColumn, AlternatingLeastSquares.java:132-136).
"""
import numpy as np


def block_rows(n, world):
    return (n + world - 1) // world


def local_block(n, rank, world):
    """[begin, end) of the rows of an n-row factor that `rank` owns."""
    b = block_rows(n, world)
    begin = min(rank * b, n)
    return begin, min(begin + b, n)


def shard_rows(ptr, idx, val, rank, world):
    """Rows [begin, end) of a CSR as a CSR whose row_ptr starts at 0 (indices stay global)."""
    ptr = np.asarray(ptr, dtype=np.int64)
    begin, end = local_block(ptr.size - 1, rank, world)
    e0, e1 = int(ptr[begin]), int(ptr[end])
    return (ptr[begin:end + 1] - e0, np.ascontiguousarray(idx[e0:e1]),
            np.ascontiguousarray(val[e0:e1]))


def padded_rows(n, world):
    return block_rows(n, world) * world


def gather_blocks(blocks, n):
    """Inverse of the per-rank block layout: concatenate equal padded blocks, drop the padding."""
    return np.concatenate(blocks, axis=0)[:n]
