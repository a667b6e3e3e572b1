"""Oracle (test infrastructure): block-sparsity layout of SparseAttention.

Restates `SparseAttention.get_master_layout` (reference sparse_vae/core/sparse_attention.py:38-59)
with explicit integer loops instead of torch's `fill_diagonal_`, plus the block enumeration order
that the Triton block-sparse ops use for their sparse tensors (row-major `(head, row, col)` order of
the non-zero blocks; reference sparse_vae/core/sparse_matmul.py:83-91,295-302).
"""
from __future__ import annotations

import numpy as np


def window_extents(window_size: int, causal: bool):
    """(left_context, right_context) exactly as sparse_attention.py:42-44 computes them."""
    num_sides = 1 if causal else 2
    q, r = divmod(window_size, num_sides)
    left_context = q + r          # "Round up"
    right_context = window_size - left_context
    return left_context, right_context


def layout_2d(num_blocks: int, window_size: int, causal: bool = True, include_cls: bool = True) -> np.ndarray:
    """[num_blocks, num_blocks] int64 0/1 layout for one head.

    sparse_attention.py:46-57: sub-diagonals 0..left_context-1 are filled, then super-diagonals
    1..right_context-1 (note: `range(1, right_context)`, i.e. one fewer than right_context), then
    column 0 when include_cls.  A slice `[:nb, :nb]` of the master layout equals the layout built at
    size nb (sparse_attention.py:63,70), which is what this function builds directly.
    """
    n = int(num_blocks)
    layout = np.zeros((n, n), dtype=np.int64)
    left_context, right_context = window_extents(window_size, causal)
    for offset in range(left_context):
        for i in range(n - offset):
            layout[offset + i, i] = 1
    for offset in range(1, right_context):
        for i in range(n - offset):
            layout[i, offset + i] = 1
    if include_cls and n > 0:
        layout[:, 0] = 1
    return layout


def master_layout(num_blocks: int, window_size: int, causal: bool = True, include_cls: bool = True,
                  num_heads: int = 8) -> np.ndarray:
    """[num_heads, num_blocks, num_blocks] int64 (sparse_attention.py:59)."""
    one = layout_2d(num_blocks, window_size, causal, include_cls)
    return np.repeat(one[None], num_heads, axis=0)


def csr(layout2d: np.ndarray):
    """(row_ptr[nb+1], col_idx[nnz]) in `layout.nonzero()` (row-major) order."""
    nb = layout2d.shape[0]
    row_ptr = np.zeros(nb + 1, dtype=np.int32)
    cols = []
    for r in range(nb):
        for c in range(nb):
            if layout2d[r, c]:
                cols.append(c)
        row_ptr[r + 1] = len(cols)
    return row_ptr, np.asarray(cols, dtype=np.int32)


def csc(layout2d: np.ndarray):
    """(colT_ptr[nb+1], rowT_idx[nnz]): for every key block the query blocks that attend to it."""
    nb = layout2d.shape[0]
    col_ptr = np.zeros(nb + 1, dtype=np.int32)
    rows = []
    for c in range(nb):
        for r in range(nb):
            if layout2d[r, c]:
                rows.append(r)
        col_ptr[c + 1] = len(rows)
    return col_ptr, np.asarray(rows, dtype=np.int32)


def nnz_closed_form(num_blocks: int, window_size: int) -> int:
    """Blocks per head for causal + include_cls, nb >= w: w(w+1)/2 + (nb-w)(w+1)."""
    w, nb = window_size, num_blocks
    return w * (w + 1) // 2 + (nb - w) * (w + 1)
