"""Tile visiting orders for GEMMs whose A rows arrive (AllGather + GEMM) or whose C rows leave (GEMM + ReduceScatter) rank by rank.

Reference: kernels/nvidia/ag_gemm_threadblock_swizzle.py:52 and gemm_rs_threadblock_swizzle.py:69 compute, inside a Triton kernel and one
warp wide, the M-tile a CTA should work on as a function of its launch index so that

* AG + GEMM starts with the rows that are already here (this rank's shard), continues with the rest of this node (NVLink) and only then
  walks the other nodes in ring order (their shards arrive last, over the NIC);
* GEMM + RS starts with the rows the ring needs first (rank + 1's) and finishes with this rank's own rows, node by node starting at
  node + 1.

Here the same orders are plain functions that return the WHOLE permutation (a table is what a persistent kernel wants: ``order[i]`` is one
global load, computed once per shape on the host), with the rule for tiles that straddle two nodes made explicit instead of patched
in by lane:  AG -- a tile that needs rows of two nodes is visited with the LATER of the two (both shards must have arrived);
RS -- it is visited with the EARLIER one (its rows must be ready when the first consumer asks).  On one node (``nnodes == 1``) the
orders are the rotations the fused kernels of this framework apply in-kernel (csrc/gemm_sm100.cuh: AG starts at the local shard, RS at
rank + 1); the multi-node tables are what a NVL72-plus-IB deployment would feed them.

``threadblock_swizzle_allgather_gemm_kernel`` / ``threadblock_swizzle_gemm_reduce_scatter_kernel`` keep the reference's call shape
(launch index in, tile index out) on top of the cached tables.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np


def _node_tile_sets(M: int, nnodes: int, block_m: int, node_order, straddle: str):
    """Tiles of every node in visiting order.  A tile whose rows belong to two (or more) nodes goes to the one visited last
    (``straddle="late"``) or first (``"early"``)."""
    n_tiles = (M + block_m - 1) // block_m
    rows_per_node = M // nnodes
    visit_pos = {n: i for i, n in enumerate(node_order)}
    sets = {n: [] for n in node_order}
    for t in range(n_tiles):
        r0, r1 = t * block_m, min(M, (t + 1) * block_m) - 1
        nodes = range(min(nnodes - 1, r0 // rows_per_node), min(nnodes - 1, r1 // rows_per_node) + 1)
        pick = (max if straddle == "late" else min)(nodes, key=lambda n: visit_pos[n])
        sets[pick].append(t)
    return [sets[n] for n in node_order]


def _rotate_to(tiles, first_tile: int, fallback_tile: int = -1):
    """Rotate an ascending tile list so that it starts at the first tile >= ``first_tile``; if there is none, at ``fallback_tile`` when
    the list has it, else at its start."""
    if not tiles:
        return tiles
    k = next((i for i, t in enumerate(tiles) if t >= first_tile), None)
    if k is None:
        k = tiles.index(fallback_tile) if fallback_tile in tiles else 0
    return tiles[k:] + tiles[:k]


@lru_cache(maxsize=256)
def allgather_gemm_tile_order(M: int, rank: int, world_size: int, nnodes: int = 1, block_m: int = 128) -> np.ndarray:
    """``order[i]`` = M tile for launch index i.  Own node first, inside a node start at the tile holding this local rank's first row."""
    assert world_size % nnodes == 0 and M % world_size == 0
    lw = world_size // nnodes
    node, local = rank // lw, rank % lw
    m_rank, m_node = M // world_size, M // nnodes
    node_order = [(node + i) % nnodes for i in range(nnodes)]
    out = []
    for n, tiles in zip(node_order, _node_tile_sets(M, nnodes, block_m, node_order, "late")):
        start_row = m_node * n + m_rank * local
        # first tile that lies entirely at / after the start of the local rank's shard (else the tile that holds that row)
        out += _rotate_to(tiles, -(-start_row // block_m), start_row // block_m)
    return np.asarray(out, dtype=np.int32)


@lru_cache(maxsize=256)
def gemm_reduce_scatter_tile_order(M: int, rank: int, world_size: int, nnodes: int = 1, block_m: int = 128) -> np.ndarray:
    """``order[i]`` = M tile for launch index i.  Node + 1 first and the own node last; inside a node start at local rank + 1's rows,
    so this rank's own rows -- which nobody waits for -- are produced last."""
    assert world_size % nnodes == 0 and M % world_size == 0
    lw = world_size // nnodes
    node, local = rank // lw, rank % lw
    m_rank, m_node = M // world_size, M // nnodes
    node_order = [(node + 1 + i) % nnodes for i in range(nnodes)]
    out = []
    for n, tiles in zip(node_order, _node_tile_sets(M, nnodes, block_m, node_order, "early")):
        start_row = m_node * n + m_rank * ((local + 1) % lw)
        out += _rotate_to(tiles, start_row // block_m)
    return np.asarray(out, dtype=np.int32)


def threadblock_swizzle_allgather_gemm_kernel(tiled_m: int, M: int, rank: int, WORLD_SIZE: int, NNODES: int, BLOCK_SIZE_M: int, DEBUG: bool = False) -> int:
    """Reference call shape (ag_gemm_threadblock_swizzle.py:52): launch index -> M tile."""
    return int(allgather_gemm_tile_order(M, rank, WORLD_SIZE, NNODES, BLOCK_SIZE_M)[tiled_m])


def threadblock_swizzle_gemm_reduce_scatter_kernel(tiled_m: int, M: int, rank: int, WORLD_SIZE: int, NNODES: int, BLOCK_SIZE_M: int, DEBUG: bool = False) -> int:
    """Reference call shape (gemm_rs_threadblock_swizzle.py:69): launch index -> M tile."""
    return int(gemm_reduce_scatter_tile_order(M, rank, WORLD_SIZE, NNODES, BLOCK_SIZE_M)[tiled_m])


threadblock_swizzle_allgather_gemm = threadblock_swizzle_allgather_gemm_kernel
threadblock_swizzle_gemm_reduce_scatter = threadblock_swizzle_gemm_reduce_scatter_kernel


def tile_order_table(order: np.ndarray, device=None):
    """int32 device tensor of a tile order (what a persistent kernel indexes with its launch index)."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(order)).to(device or "cpu")


# ---- AllGather + grouped GEMM (TP-MoE up projection): tiles ordered by the arrival of the tokens they need ------------------------------
def ag_moe_tile_table(ntokens_per_rank_per_expert, rank: int, block_m: int = 128) -> np.ndarray:
    """Tile table of the grouped GEMM behind an all-gather of tokens, in EXECUTION order.

    Reference: kernels/nvidia/threadblock_swizzle_ag_moe{.py,_triton.py,.cu,.cc} (N12): rows are sorted by (expert, arrival stage of the
    source rank) -- stage of source s for this rank is ``(s - rank) % world``, the own shard is stage 0 -- so a 128-row tile of expert e
    needs the shards of the sources between the first and the last row it contains; tiles whose last needed shard arrives earlier run
    earlier.  Input: ``[world, E]`` token counts (pairs routed from every source rank to every expert).  Output: int32 ``[n_tiles, 4]``
    rows ``(expert, tile index inside the expert, first stage, last stage)`` ordered by (last stage, expert, tile); a consumer waits for
    the flags of stages ``first..last`` (``moe_align_sort(..., tokens_per_rank=, rank=, world=)`` produces the matching row order)."""
    cnt = np.asarray(ntokens_per_rank_per_expert, dtype=np.int64)
    world, E = cnt.shape
    by_stage = np.stack([cnt[(rank + st) % world] for st in range(world)])            # [stage, E]
    tiles = []
    for e in range(E):
        ends = np.cumsum(by_stage[:, e])                                               # rows of expert e up to and including every stage
        total = int(ends[-1])
        for t in range((total + block_m - 1) // block_m):
            r0, r1 = t * block_m, min(total, (t + 1) * block_m) - 1
            first = int(np.searchsorted(ends, r0, side="right"))
            last = int(np.searchsorted(ends, r1, side="right"))
            tiles.append((e, t, first, last))
    tiles.sort(key=lambda x: (x[3], x[0], x[1]))
    return np.asarray(tiles, dtype=np.int32).reshape(-1, 4)


def check_ag_moe_tile_table(table: np.ndarray, ntokens_per_rank_per_expert, rank: int, block_m: int = 128) -> bool:
    """Every tile of every expert appears exactly once, its stage range covers exactly the sources of its rows, and the execution order
    never schedules a tile before one that needs strictly earlier shards only."""
    cnt = np.asarray(ntokens_per_rank_per_expert, dtype=np.int64)
    world, E = cnt.shape
    want = {(e, t) for e in range(E) for t in range((int(cnt[:, e].sum()) + block_m - 1) // block_m)}
    got = [(int(r[0]), int(r[1])) for r in table]
    if len(got) != len(set(got)) or set(got) != want:
        return False
    for e, t, first, last in table.tolist():
        stage_of_row = np.repeat(np.arange(world), [int(cnt[(rank + st) % world, e]) for st in range(world)])
        rows = stage_of_row[t * block_m:(t + 1) * block_m]
        if rows.size == 0 or int(rows[0]) != first or int(rows[-1]) != last:
            return False
    lasts = table[:, 3]
    return bool(np.all(lasts[:-1] <= lasts[1:]))


threadblock_swizzle_ag_moe = ag_moe_tile_table
check_swizzled = check_ag_moe_tile_table

