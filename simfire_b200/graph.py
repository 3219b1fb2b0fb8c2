"""
Fire-spread graph, rebuilt after the fact from the ignition-step plane.

The reference keeps a networkx DiGraph with one node per pixel and, every update, adds an edge
to each newly burning pixel from every adjacent pixel (always all 8, whatever
`diagonal_spread` says) whose fire_map value is BURNING at that moment
(`FireSpreadGraph.add_edges_from_manager`, simfire/utils/graph.py:84-150, called from
`_update_with_new_locs`, fire.py:582-584) -- 8 % of its step time and 0.6 GB at 1024 x 1024.
Nothing in the step reads the graph, so here it is not maintained per step: the device keeps one
int32 per cell, the update() call that ignited it (`keep_ignition=True`), and the edges follow
from it: a pixel ignited by call s is BURNING as seen by calls s+1 .. s+max_fire_duration
(pruned at the start of call s+max_fire_duration+1, fire.py:116-161), so

    edge (n -> d)   iff   n adjacent to d,  ign[n] >= 0,  ign[n] < ign[d] <= ign[n] + max_fire_duration.

Exact as long as fire_map cells are not overwritten while BURNING (mitigation drawn over live
fire, `load_mitigation` with BURNING cells); those edits are not part of the ignition history.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

# adjacency order of graph.py:121-130 as (dx, dy)
ADJACENT = ((1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1))


def spread_edges(ignition: np.ndarray, max_fire_duration: int) -> np.ndarray:
    """
    Edges of the fire-spread graph as an int32 array [n_edges, 4] of (x_src, y_src, x_dst, y_dst),
    from an (H, W) ignition-step plane (-1 = never ignited, 0 = initial fire).
    """
    ign = np.asarray(ignition)
    H, W = ign.shape
    out = []
    dst_ok = ign > 0  # the initial fire has no parent
    for dx, dy in ADJACENT:
        # source n = d + (dx, dy): shift the plane so that src[y, x] = ign[y + dy, x + dx]
        src = np.full((H, W), -1, dtype=ign.dtype)
        ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
        xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
        src[yd, xd] = ign[ys, xs]
        hit = dst_ok & (src >= 0) & (src < ign) & (ign <= src + max_fire_duration)
        y, x = np.nonzero(hit)
        out.append(np.stack([x + dx, y + dy, x, y], axis=1))
    return np.concatenate(out).astype(np.int32) if out else np.zeros((0, 4), np.int32)


def to_networkx(ignition: np.ndarray, max_fire_duration: int):
    """The reference's graph object: a DiGraph over all (x, y) pixels (graph.py:18-51)."""
    import networkx as nx

    H, W = np.asarray(ignition).shape
    g = nx.DiGraph()
    g.add_nodes_from((x, y) for y in range(H) for x in range(W))
    e = spread_edges(ignition, max_fire_duration)
    g.add_edges_from(((int(a), int(b)), (int(c), int(d))) for a, b, c, d in e)
    return g


def edge_set(edges: np.ndarray) -> set:
    return {((int(a), int(b)), (int(c), int(d))) for a, b, c, d in np.asarray(edges).reshape(-1, 4)}


__all__ = ["spread_edges", "to_networkx", "edge_set", "ADJACENT"]
