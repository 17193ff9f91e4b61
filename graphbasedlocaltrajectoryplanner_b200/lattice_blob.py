"""
Packs a ``Lattice`` into ONE contiguous byte blob + ``LatticeHeader`` (include/ltpl_b200.h) and derives the fixed
capacities (``Dims``) of the batched kernels from it.  A single blob is what rank 0 broadcasts over NCCL to the other
ranks; every section is 256-byte aligned so that 128-bit loads are always aligned.
"""

from __future__ import annotations

import bisect

import numpy as np

from . import capi
from .lattice import Lattice

ALIGN = 256


def end_layer_of(lat: Lattice, start_layer: int) -> tuple:
    """planning range end layer and layer distance (reference: gen_local_node_template.py:104-142)."""
    if lat.plan_horizon_mode == 'distance':
        des = lat.s_raceline[start_layer] + lat.min_plan_horizon
        if des > lat.s_raceline[-1]:
            if lat.closed:
                des -= lat.s_raceline[-1]
            else:
                des = lat.s_raceline[-1]
        end = bisect.bisect_left(lat.s_raceline, des)
    elif lat.plan_horizon_mode == 'layers':
        if lat.closed:
            end = (start_layer + int(lat.min_plan_horizon)) % lat.num_layers
        else:
            end = max(start_layer + int(lat.min_plan_horizon), lat.num_layers - 1)
    else:
        raise ValueError('Unsupported planning horizon mode "' + str(lat.plan_horizon_mode) + '"!')
    dist = end - start_layer
    if dist < 0:
        dist = lat.num_layers - start_layer + end
    return int(end), int(dist)


def capacities(lat: Lattice) -> dict:
    """worst-case sizes over all possible start layers."""
    L = lat.num_layers
    n_edges_pair = np.diff(lat.edge_layer_off).astype(np.int64)
    nsamp = np.diff(lat.samp_off).astype(np.int64)
    sl = lat.edge_start_layer()
    max_samp_pair = np.zeros(L, dtype=np.int64)
    if nsamp.size:
        np.maximum.at(max_samp_pair, sl, nsamp)
    max_dist, max_we, max_pnew = 0, 0, 0
    for s in range(L):
        _, dist = end_layer_of(lat, s)
        pairs = [(s + k) % L for k in range(dist)]
        max_dist = max(max_dist, dist)
        max_we = max(max_we, int(n_edges_pair[pairs].sum()) if pairs else 0)
        max_pnew = max(max_pnew, int(np.maximum(max_samp_pair[pairs] - 1, 0).sum()) + 1 if pairs else 1)
    # constant segment: pose -> race line node of layer (l + 2) % (L - 1) where l is the layer of the closest node
    # (OTH:223-229); the pose can sit anywhere between the neighbouring layers
    xy = np.column_stack((lat.node_x, lat.node_y))
    worst = 0.0
    for l in range(L):
        g_layer = (l + 2) % (L - 1)
        goal = xy[lat.node_off[g_layer] + lat.raceline_index[g_layer]]
        for ll in ((l - 1) % L, l, (l + 1) % L):
            pts = xy[lat.node_off[ll]:lat.node_off[ll + 1]]
            if pts.size:
                worst = max(worst, float(np.sqrt(((pts - goal) ** 2).sum(axis=1)).max()))
    p0_max = int(np.ceil(1.25 * worst / lat.sampled_resolution)) + 4
    p_max = ((p0_max + max_pnew + 3) // 4) * 4
    return dict(h_max=max_dist + 2, max_window_edges=max_we, p0_max=p0_max, p_max=p_max,
                max_plan_layers=max_dist)


GRID_CELL = 4.0        # m
GRID_REACH = 60.0      # m: cells whose centre is further from the polyline get no bound (whole-polyline scan)
GRID_MAX_COUNT = 32    # one vertex per lane


def nearest_grid(pts: np.ndarray, closed: bool, x0: float, y0: float, nx: int, ny: int, cell: float = GRID_CELL) -> np.ndarray:
    """int32 [ny][nx]: entry = first << 6 | count.  For EVERY position q inside cell (ix, iy) the nearest vertex of
    ``pts`` -- and every vertex at the same distance, so np.argmin's first-minimum rule survives -- lies among the
    ``count`` vertices first, first + 1, ... (indices modulo len(pts) when ``closed``).  Proof: with m the cell centre and
    h its half diagonal, |d(q, v) - d(m, v)| <= h for all v, hence d(m, v*) <= min_v d(m, v) + 2 h for the nearest vertex
    v* of q.  count = 0: no bound (far from the polyline, or the candidates span more than 32 indices)."""
    n = pts.shape[0]
    h = cell * np.sqrt(0.5)
    cx = x0 + (np.arange(nx) + 0.5) * cell
    cy = y0 + (np.arange(ny) + 0.5) * cell
    out = np.zeros((ny, nx), dtype=np.int32)
    for iy in range(ny):
        dx = cx[:, None] - pts[None, :, 0]
        dy = cy[iy] - pts[None, :, 1]
        dist = np.sqrt(dx * dx + dy * dy)                       # (nx, n)
        dmin = dist.min(axis=1)
        for ix in np.nonzero(dmin <= GRID_REACH)[0]:
            idx = np.nonzero(dist[ix] <= dmin[ix] + 2.0 * h + 1e-3)[0]
            if closed and idx.size > 1:
                gaps = np.diff(np.append(idx, idx[0] + n))      # cyclic distance to the next candidate
                k = int(np.argmax(gaps))
                first, count = int(idx[(k + 1) % idx.size]), int(n - gaps[k] + 1)
            else:
                first, count = int(idx[0]), int(idx[-1] - idx[0] + 1)
            if count <= GRID_MAX_COUNT:
                out[iy, ix] = (first << 6) | count
    return out


def pack_lattice(lat: Lattice) -> tuple:
    """returns (LatticeHeader, blob uint8 ndarray, capacities dict)."""
    cap = capacities(lat)
    if lat.max_nodes_per_layer > 64:
        raise ValueError("lattices with more than 64 nodes per layer are not supported by the DP kernel")
    if not lat.virt_goal_node:
        raise NotImplementedError("virt_goal_n=False (sequential goal-node probing, GB:896-927) is not batched")
    bound1 = lat.refline + lat.normvec * np.expand_dims(lat.w_right, 1)      # OTH:208-211, OLI:70-72
    bound2 = lat.refline - lat.normvec * np.expand_dims(lat.w_left, 1)
    center = (bound1 + bound2) / 2                                          # check_inside_bounds.py:27
    nn = lat.num_nodes
    node_layer = np.repeat(np.arange(lat.num_layers, dtype=np.int32), np.diff(lat.node_off))
    nsamp = np.diff(lat.samp_off)
    samp_edge = np.repeat(np.arange(lat.num_edges, dtype=np.int32), nsamp)
    g = lat.glob_rl
    glob6 = np.column_stack((g[:-1], np.diff(g[:, 0])))                     # CVPF:166
    edge_rec = np.zeros(lat.num_edges, dtype=np.dtype([("cost", "<f8"), ("src", "<i4"), ("dst", "<i4")]))
    edge_rec["cost"], edge_rec["src"], edge_rec["dst"] = lat.edge_cost, lat.edge_src, lat.edge_dst
    tab_stride = int(cap["h_max"])
    # the device back-track identifies an edge by its (source node, destination node): one edge per node pair
    e_layer = np.repeat(np.arange(lat.num_layers, dtype=np.int64), np.diff(lat.edge_layer_off))
    key = (e_layer * 64 + lat.edge_src.astype(np.int64)) * 64 + lat.edge_dst.astype(np.int64)
    if np.unique(key).size != key.size:
        raise ValueError("lattice holds parallel edges between one pair of nodes")
    # nearest-vertex grids of the four polylines the online path searches for every object (cached on the lattice)
    grids = getattr(lat, "_nearest_grids", None)
    if grids is None:
        allp = np.vstack((center, lat.refline, lat.raceline, glob6[:, 1:3]))
        gx0, gy0 = (allp.min(axis=0) - GRID_REACH)
        gnx, gny = (int(v) for v in np.ceil((allp.max(axis=0) + GRID_REACH - (gx0, gy0)) / GRID_CELL))
        closed = bool(lat.closed)
        grids = dict(x0=float(gx0), y0=float(gy0), nx=gnx, ny=gny,
                     center=nearest_grid(center, closed, gx0, gy0, gnx, gny),
                     refline=nearest_grid(lat.refline, closed, gx0, gy0, gnx, gny),
                     raceline=nearest_grid(lat.raceline, closed, gx0, gy0, gnx, gny),
                     glob=nearest_grid(np.ascontiguousarray(glob6[:, 1:3]), closed, gx0, gy0, gnx, gny))
        try:
            lat._nearest_grids = grids
        except AttributeError:
            pass

    sections = [
        ("off_node_off", lat.node_off.astype(np.int32)),
        ("off_raceline_index", lat.raceline_index.astype(np.int32)),
        ("off_s_raceline", lat.s_raceline.astype(np.float64)),
        ("off_vel_raceline", lat.vel_raceline.astype(np.float64)),
        ("off_refline", lat.refline.astype(np.float64)),
        ("off_raceline", lat.raceline.astype(np.float64)),
        ("off_bound1", bound1.astype(np.float64)),
        ("off_bound2", bound2.astype(np.float64)),
        ("off_centerline", center.astype(np.float64)),
        ("off_node_xy", np.column_stack((lat.node_x, lat.node_y)).astype(np.float64)),
        ("off_node_psi", lat.node_psi.astype(np.float64)),
        ("off_node_layer", node_layer),
        ("off_in_off", lat.in_off.astype(np.int32).reshape(nn, 2)),
        ("off_edge_layer_off", lat.edge_layer_off.astype(np.int32)),
        ("off_edge_src", lat.edge_src.astype(np.int32)),
        ("off_edge_dst", lat.edge_dst.astype(np.int32)),
        ("off_edge_cost", lat.edge_cost.astype(np.float64)),
        ("off_edge_len", lat.edge_len.astype(np.float64)),
        ("off_edge_psi1", lat.edge_psi1.astype(np.float64)),
        ("off_edge_psi0", lat.edge_psi0.astype(np.float64)),
        ("off_samp_off", lat.samp_off.astype(np.int32)),
        ("off_samp_xy", np.column_stack((lat.samp_x, lat.samp_y)).astype(np.float64)),
        ("off_samp_el", lat.samp_el.astype(np.float64)),
        ("off_samp_edge", samp_edge),
        ("off_glob_rl", glob6.astype(np.float64)),
        ("off_glob_xy", np.ascontiguousarray(glob6[:, 1:3]).astype(np.float64)),
        ("off_edge_rec", edge_rec),
        # follow table: zero here, filled on the device by ltpl_lattice_create (k_follow_table)
        ("off_tab_reach", np.zeros(nn, dtype=np.int32)),
        ("off_tab_node", np.zeros(nn * tab_stride, dtype=np.uint8)),
        ("off_tab_edge", np.zeros(nn * tab_stride, dtype=np.int32)),
        ("off_grid_center", grids["center"]),
        ("off_grid_refline", grids["refline"]),
        ("off_grid_raceline", grids["raceline"]),
        ("off_grid_glob", grids["glob"]),
    ]
    h = capi.LatticeHeader()
    h.abi_version = capi.ABI_VERSION
    h.num_layers, h.num_nodes, h.num_edges, h.num_samples = lat.num_layers, nn, lat.num_edges, lat.num_samples
    h.n_glob_rl = int(g.shape[0])
    h.closed = int(bool(lat.closed))
    h.plan_horizon_mode = 0 if lat.plan_horizon_mode == 'distance' else 1
    if lat.plan_horizon_mode not in ('distance', 'layers'):
        raise ValueError('Unsupported planning horizon mode "' + str(lat.plan_horizon_mode) + '"!')
    h.max_nodes_per_layer = lat.max_nodes_per_layer
    h.max_window_edges = cap["max_window_edges"]
    h.max_pair_edges = max(int(np.diff(lat.edge_layer_off).max()), 1)
    h.tab_stride = tab_stride
    h.grid_nx, h.grid_ny, h.grid_x0, h.grid_y0, h.grid_inv_cell = grids["nx"], grids["ny"], grids["x0"], grids["y0"], 1.0 / GRID_CELL
    h.lat_offset, h.lat_resolution, h.sampled_resolution = lat.lat_offset, lat.lat_resolution, lat.sampled_resolution
    h.vel_decrease_lat, h.veh_width, h.veh_length = lat.vel_decrease_lat, lat.veh_width, lat.veh_length
    h.virt_goal_node_cost, h.min_plan_horizon = lat.virt_goal_node_cost, lat.min_plan_horizon

    off = 0
    chunks = []
    for name, arr in sections:
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        setattr(h, name, off)
        pad = (-raw.size) % ALIGN
        chunks.append(raw)
        if pad:
            chunks.append(np.zeros(pad, dtype=np.uint8))
        off += raw.size + pad
    h.blob_bytes = off
    blob = np.concatenate(chunks)
    assert blob.size == off
    return h, blob, cap
