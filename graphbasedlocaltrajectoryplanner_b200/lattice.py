"""
Flat SoA lattice ("lattice blob") -- the input of the batched online planning path.

The reference keeps the offline graph as a pickled ``GraphBase`` wrapping two ``igraph.Graph`` copies with Python
objects as vertex / edge attributes (/root/reference/graph_ltpl/data_objects/GraphBase.py:93-135, 163-194, 409-439).
The online path only ever reads: node positions / headings, per-layer race line bookkeeping, and per edge the sampled
points (x, y, psi, kappa, chord length), the chord-length sum and the offline cost (MOPG:274-295, GB:626-644, 818-821).
This module holds exactly that, flattened for coalesced device access:

  per layer  l in [0, L):   node_off[l] .. node_off[l+1]      global node ids of layer l
                            raceline_index, s_raceline, vel_raceline, refline, normvec, w_right, w_left, alpha, raceline
  per node   g:             node_x, node_y, node_psi;  in_off[g] = (first in-edge, #in-edges) (all from the previous layer)
  per edge   e (CSC, sorted by (start_layer, dst_node, src_node)):
                            edge_src (local node id in the start layer), edge_dst (local id in the end layer),
                            edge_cost (float64 offline cost), edge_len (sum of chord lengths), edge_psi0 / edge_psi1
                            (heading at first / last sample), samp_off[e] .. samp_off[e+1] = samples
  per sample s:             samp_x, samp_y (float64: collision decisions must be bit-exact), samp_psi, samp_kappa,
                            samp_el (chord length to the next sample, trailing 0 -- GB:425-436)

Two producers:
  * ``Lattice.from_graph_base(gb)``  -- extractor from a live reference ``GraphBase`` through its public API
    (get_nodes GB:359, get_edges GB:648, get_edge GB:444, get_node_info GB:221); honours "offline_graph pickle as
    input" whenever the reference + igraph are importable.
  * ``build_lattice(track_csv, offline_ini)`` -- NumPy restatement of the reference's offline pipeline
    (imp_global_traj/src/import_globtraj_csv.py:34-58, variable_step_size.py:31-56,
    offline_graph/src/main_offline_callback.py:76-179, gen_node_skeleton.py:43-166, gen_edges.py:34-163,
    prune_graph.py:24-67, gen_offline_cost.py:47-79) for boxes where neither igraph nor tph exist (the GPU box).
"""

from __future__ import annotations

import configparser
import hashlib
import math
import os
from dataclasses import dataclass, field

import numpy as np

LATTICE_FORMAT_VERSION = 3


# ----------------------------------------------------------------------------------------------------------------------
# small helpers (own implementations of the tph pieces the offline pipeline needs)
# ----------------------------------------------------------------------------------------------------------------------
def _normalize_psi(psi: np.ndarray) -> np.ndarray:
    out = np.sign(psi) * np.mod(np.abs(psi), 2 * math.pi)
    out = np.array(out, dtype=np.float64, copy=True)
    out[out >= math.pi] -= 2 * math.pi
    out[out < -math.pi] += 2 * math.pi
    return out


def _heading_closed(path: np.ndarray) -> np.ndarray:
    """Secant heading of a closed polyline with +-1 index preview/review (0 = north).

    This is what the reference obtains from tph.calc_head_curv_num at gen_node_skeleton.py:63-65,86-92: with layer
    spacings of 6..30 m the 1 m preview/review distances round to 0 index steps and are clamped to 1.
    """
    nxt = np.roll(path, -1, axis=0)
    prv = np.roll(path, 1, axis=0)
    tang = nxt - prv
    return _normalize_psi(np.arctan2(tang[:, 1], tang[:, 0]) - math.pi / 2)


def _closed_spline_coeffs(path_cl: np.ndarray) -> tuple:
    """C2 closed cubic spline through path_cl (first point repeated at the end) with distance scaling
    (tph.calc_splines closed branch as used at gen_edges.py:46-47).  Dense solve; runs once per lattice."""
    n = path_cl.shape[0] - 1
    el = np.sqrt(np.sum(np.diff(path_cl, axis=0) ** 2, axis=1))
    el = np.append(el, el[0])
    scaling = el[:-1] / el[1:]

    m = np.zeros((4 * n, 4 * n))
    bx = np.zeros(4 * n)
    by = np.zeros(4 * n)
    for i in range(n):
        j = 4 * i
        m[j, j] = 1.0
        m[j + 1, j:j + 4] = 1.0
        bx[j], bx[j + 1] = path_cl[i, 0], path_cl[i + 1, 0]
        by[j], by[j + 1] = path_cl[i, 1], path_cl[i + 1, 1]
        if i < n - 1:
            m[j + 2, j + 1:j + 4] = (1.0, 2.0, 3.0)
            m[j + 2, j + 5] = -scaling[i]
            m[j + 3, j + 2:j + 4] = (2.0, 6.0)
            m[j + 3, j + 6] = -2.0 * scaling[i] ** 2
    # periodic heading / curvature rows
    m[-2, 1] = scaling[-1]
    m[-2, -3:] = (-1.0, -2.0, -3.0)
    m[-1, 2] = 2.0 * scaling[-1] ** 2
    m[-1, -2:] = (-2.0, -6.0)
    cx = np.linalg.solve(m, bx).reshape(n, 4)
    cy = np.linalg.solve(m, by).reshape(n, 4)
    return cx, cy


_M2 = np.array([[1.0, 0.0, 0.0, 0.0],
                [1.0, 1.0, 1.0, 1.0],
                [0.0, 1.0, 0.0, 0.0],
                [0.0, 1.0, 2.0, 3.0]])


def _two_point_splines(p0: np.ndarray, p1: np.ndarray, psi_s: np.ndarray, psi_e: np.ndarray) -> tuple:
    """Batched single-segment cubic with heading boundary conditions (tph.calc_splines with N=1, gen_edges.py:88-92).
    Rows: a0 = P0, a0+a1+a2+a3 = P1, a1 = cos(psi_s+pi/2) el, a1+2a2+3a3 = cos(psi_e+pi/2) el (el = |P1-P0|)."""
    el = np.sqrt(np.sum((p1 - p0) ** 2, axis=1))
    n = p0.shape[0]
    bx = np.stack((p0[:, 0], p1[:, 0], np.cos(psi_s + math.pi / 2) * el, np.cos(psi_e + math.pi / 2) * el), axis=1)
    by = np.stack((p0[:, 1], p1[:, 1], np.sin(psi_s + math.pi / 2) * el, np.sin(psi_e + math.pi / 2) * el), axis=1)
    mm = np.broadcast_to(_M2, (n, 4, 4))
    cx = np.linalg.solve(mm, bx[:, :, None])[:, :, 0]
    cy = np.linalg.solve(mm, by[:, :, None])[:, :, 0]
    return cx, cy


def variable_step_size(kappa, dist, d_curve, d_straight, curve_th, force_last=False) -> list:
    """Curvature dependent layer spacing (reference: imp_global_traj/src/variable_step_size.py:31-56)."""
    next_dist = 0.0
    next_dist_min = 0.0
    cur_dist = 0.0
    idx_array = []
    for idx in range(min(len(kappa), len(dist))):
        d = dist[idx]
        if (cur_dist + d) > next_dist_min and abs(kappa[idx]) > curve_th:
            next_dist = cur_dist
        if (cur_dist + d) > next_dist:
            idx_array.append(idx)
            if abs(kappa[idx]) < curve_th:
                next_dist += d_straight
            else:
                next_dist += d_curve
            next_dist_min = cur_dist + d_curve
        cur_dist += d
    if force_last and len(kappa) - 1 not in idx_array:
        idx_array.append(len(kappa) - 1)
    return idx_array


def import_globtraj_csv(import_path: str) -> dict:
    """12-column ';'-separated global trajectory file (reference: import_globtraj_csv.py:34-58)."""
    d = np.loadtxt(import_path, delimiter=';')
    return dict(refline=d[:-1, 0:2], width_right=d[:-1, 2], width_left=d[:-1, 3], normvec=d[:-1, 4:6],
                alpha=d[:-1, 6], length_rl=np.diff(d[:, 7]), kappa_rl=d[:-1, 9], vel_rl=d[:-1, 10],
                s_rl=d[:, 7], psi_rl=d[:-1, 8])


def md5_of_files(*paths) -> str:
    out = ""
    for p in paths:
        h = hashlib.md5()
        with open(p, "rb") as fh:
            for chunk in iter(lambda: fh.read(4096), b""):
                h.update(chunk)
        out += h.hexdigest()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# container
# ----------------------------------------------------------------------------------------------------------------------
_ARRAY_FIELDS = ("node_off", "raceline_index", "s_raceline", "vel_raceline", "refline", "normvec", "w_right", "w_left",
                 "alpha", "raceline", "node_x", "node_y", "node_psi", "in_off", "edge_layer_off", "edge_src",
                 "edge_dst", "edge_cost", "edge_len", "edge_psi0", "edge_psi1", "samp_off", "samp_x", "samp_y",
                 "samp_psi", "samp_kappa", "samp_el", "glob_rl")
_SCALAR_FIELDS = ("num_layers", "closed", "lat_offset", "lat_resolution", "sampled_resolution", "vel_decrease_lat",
                  "veh_width", "veh_length", "veh_turn", "virt_goal_node", "virt_goal_node_cost", "min_plan_horizon",
                  "plan_horizon_mode", "md5_params")


@dataclass
class Lattice:
    # scalars (GraphBase ctor, GB:93-116)
    num_layers: int
    closed: bool
    lat_offset: float
    lat_resolution: float
    sampled_resolution: float
    vel_decrease_lat: float
    veh_width: float
    veh_length: float
    veh_turn: float
    virt_goal_node: bool
    virt_goal_node_cost: float
    min_plan_horizon: float
    plan_horizon_mode: str
    md5_params: str
    # per layer
    node_off: np.ndarray
    raceline_index: np.ndarray
    s_raceline: np.ndarray
    vel_raceline: np.ndarray
    refline: np.ndarray
    normvec: np.ndarray
    w_right: np.ndarray
    w_left: np.ndarray
    alpha: np.ndarray
    raceline: np.ndarray
    # per node
    node_x: np.ndarray
    node_y: np.ndarray
    node_psi: np.ndarray
    in_off: np.ndarray
    # per edge
    edge_layer_off: np.ndarray
    edge_src: np.ndarray
    edge_dst: np.ndarray
    edge_cost: np.ndarray
    edge_len: np.ndarray
    edge_psi0: np.ndarray
    edge_psi1: np.ndarray
    samp_off: np.ndarray
    # per sample
    samp_x: np.ndarray
    samp_y: np.ndarray
    samp_psi: np.ndarray
    samp_kappa: np.ndarray
    samp_el: np.ndarray
    # global race line (s, x, y, kappa, vel), closed: first row repeated (main_offline_callback.py:100-104)
    glob_rl: np.ndarray
    _edge_lookup: dict = field(default=None, repr=False, compare=False)

    # -- convenience ---------------------------------------------------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        return int(self.node_off[-1])

    @property
    def num_edges(self) -> int:
        return int(self.edge_src.shape[0])

    @property
    def num_samples(self) -> int:
        return int(self.samp_x.shape[0])

    def nodes_in_layer(self, layer: int) -> int:
        return int(self.node_off[layer + 1] - self.node_off[layer])

    @property
    def max_nodes_per_layer(self) -> int:
        return int(np.max(np.diff(self.node_off)))

    def edge_start_layer(self) -> np.ndarray:
        """start layer of every edge (derived from edge_layer_off)."""
        return np.repeat(np.arange(self.num_layers), np.diff(self.edge_layer_off)).astype(np.int32)

    def edge_id(self, start_layer: int, start_node: int, end_node: int) -> int:
        """edge id of (start_layer, start_node) -> (start_layer + 1, end_node); -1 if absent (GB.get_edge, GB:444)."""
        if self._edge_lookup is None:
            sl = self.edge_start_layer()
            self._edge_lookup = {(int(a), int(b), int(c)): i
                                 for i, (a, b, c) in enumerate(zip(sl, self.edge_src, self.edge_dst))}
        return self._edge_lookup.get((int(start_layer), int(start_node), int(end_node)), -1)

    def summary(self) -> dict:
        return dict(num_layers=self.num_layers, nodes=self.num_nodes, edges=self.num_edges, samples=self.num_samples,
                    max_nodes_per_layer=self.max_nodes_per_layer,
                    max_edges_per_layer=int(np.max(np.diff(self.edge_layer_off))),
                    max_samples_per_edge=int(np.max(np.diff(self.samp_off))))

    # -- (de)serialisation -----------------------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        payload = {k: getattr(self, k) for k in _ARRAY_FIELDS}
        for k in _SCALAR_FIELDS:
            payload["scalar__" + k] = np.array(getattr(self, k))
        payload["format_version"] = np.array(LATTICE_FORMAT_VERSION)
        tmp = path + ".tmp%d.npz" % os.getpid()
        np.savez(tmp, **payload)
        os.replace(tmp, path)

    @staticmethod
    def load(path: str) -> "Lattice":
        with np.load(path, allow_pickle=False) as z:
            if int(z["format_version"]) != LATTICE_FORMAT_VERSION:
                raise ValueError("lattice blob format version mismatch")
            kw = {k: z[k] for k in _ARRAY_FIELDS}
            for k in _SCALAR_FIELDS:
                v = z["scalar__" + k]
                kw[k] = v.item()
        return Lattice(**kw)

    # -- producer (a): extractor from a live reference GraphBase ------------------------------------------------------------
    @staticmethod
    def from_graph_base(gb) -> "Lattice":
        """Flatten a reference ``GraphBase`` through its public API only (GB:221, 359, 444, 648)."""
        num_layers = int(gb.num_layers)
        nil = [int(gb.nodes_in_layer[l]) for l in range(num_layers)]
        node_off = np.concatenate(([0], np.cumsum(nil))).astype(np.int32)
        nn = int(node_off[-1])
        node_x = np.zeros(nn)
        node_y = np.zeros(nn)
        node_psi = np.zeros(nn)
        for (l, n) in gb.get_nodes():
            pos, psi, _, _, _ = gb.get_node_info(layer=l, node_number=n, active_filter=None)
            g = node_off[l] + n
            node_x[g], node_y[g], node_psi[g] = pos[0], pos[1], psi

        recs = []
        for (ls, ns, le, ne) in gb.get_edges():
            _, sp, cost, slen = gb.get_edge(ls, ns, le, ne)
            recs.append((int(ls), int(ne), int(ns), float(cost), float(slen), np.asarray(sp, dtype=np.float64)))
        return _assemble(num_layers=num_layers, closed=bool(gb.closed), node_off=node_off, node_x=node_x, node_y=node_y,
                         node_psi=node_psi, recs=recs, gb_like=gb)


def _assemble(num_layers, closed, node_off, node_x, node_y, node_psi, recs, gb_like) -> Lattice:
    """Sort edge records (start_layer, dst, src, cost, len, param(S,5)) into the CSC layout and build the SoA."""
    recs.sort(key=lambda r: (r[0], r[1], r[2]))
    ne = len(recs)
    edge_src = np.array([r[2] for r in recs], dtype=np.int32)
    edge_dst = np.array([r[1] for r in recs], dtype=np.int32)
    edge_sl = np.array([r[0] for r in recs], dtype=np.int32)
    edge_cost = np.array([r[3] for r in recs], dtype=np.float64)
    edge_len = np.array([r[4] for r in recs], dtype=np.float64)
    nsamp = np.array([r[5].shape[0] for r in recs], dtype=np.int64)
    samp_off = np.concatenate(([0], np.cumsum(nsamp))).astype(np.int32)
    allp = np.concatenate([r[5] for r in recs], axis=0) if ne else np.zeros((0, 5))
    edge_psi0 = allp[samp_off[:-1], 2].copy() if ne else np.zeros(0)
    edge_psi1 = allp[samp_off[1:] - 1, 2].copy() if ne else np.zeros(0)

    edge_layer_off = np.searchsorted(edge_sl, np.arange(num_layers + 1), side="left").astype(np.int32)
    # in-edge ranges per global destination node (destination layer = start layer + 1, wrapping on closed tracks)
    dst_layer = (edge_sl + 1) % num_layers
    gdst = node_off[dst_layer] + edge_dst
    nn = int(node_off[-1])
    cnt = np.bincount(gdst, minlength=nn)
    # edges are sorted by (start_layer, dst): the in-edges of one global node are contiguous
    first = np.zeros(nn, dtype=np.int64)
    if ne:
        chg = np.concatenate(([True], gdst[1:] != gdst[:-1]))
        first[gdst[chg]] = np.nonzero(chg)[0]
    in_start = first.astype(np.int32)
    in_cnt = cnt.astype(np.int32)

    g = gb_like
    lat = Lattice(
        num_layers=int(num_layers), closed=bool(closed), lat_offset=float(g.lat_offset),
        lat_resolution=float(g.lat_resolution), sampled_resolution=float(g.sampled_resolution),
        vel_decrease_lat=float(g.vel_decrease_lat), veh_width=float(g.veh_width), veh_length=float(g.veh_length),
        veh_turn=float(g.veh_turn), virt_goal_node=bool(g.virt_goal_node),
        virt_goal_node_cost=float(g.virt_goal_node_cost), min_plan_horizon=float(g.min_plan_horizon),
        plan_horizon_mode=str(g.plan_horizon_mode), md5_params=str(g.md5_params),
        node_off=np.asarray(node_off, dtype=np.int32),
        raceline_index=np.asarray(g.raceline_index, dtype=np.int32),
        s_raceline=np.asarray(g.s_raceline, dtype=np.float64), vel_raceline=np.asarray(g.vel_raceline, dtype=np.float64),
        refline=np.asarray(g.refline, dtype=np.float64), normvec=np.asarray(g.normvec_normalized, dtype=np.float64),
        w_right=np.asarray(g.track_width_right, dtype=np.float64),
        w_left=np.asarray(g.track_width_left, dtype=np.float64), alpha=np.asarray(g.alpha, dtype=np.float64),
        raceline=np.asarray(g.raceline, dtype=np.float64),
        node_x=node_x, node_y=node_y, node_psi=node_psi,
        in_off=np.stack((in_start, in_cnt), axis=1).astype(np.int32),
        edge_layer_off=edge_layer_off, edge_src=edge_src, edge_dst=edge_dst, edge_cost=edge_cost, edge_len=edge_len,
        edge_psi0=edge_psi0, edge_psi1=edge_psi1, samp_off=samp_off,
        samp_x=allp[:, 0].copy(), samp_y=allp[:, 1].copy(), samp_psi=allp[:, 2].copy(), samp_kappa=allp[:, 3].copy(),
        samp_el=allp[:, 4].copy(), glob_rl=np.asarray(g.glob_rl, dtype=np.float64))
    return lat


# ----------------------------------------------------------------------------------------------------------------------
# producer (b): NumPy restatement of the offline pipeline
# ----------------------------------------------------------------------------------------------------------------------
class _GraphParams(object):
    """attribute bag mirroring the public fields of GraphBase that `_assemble` reads."""
    pass


def read_offline_config(path: str) -> dict:
    cfg = configparser.ConfigParser()
    if not cfg.read(path):
        raise ValueError('Specified graph config file does not exist or is empty!')
    return dict(
        lat_resolution=cfg.getfloat('LATTICE', 'lat_resolution'),
        variable_heading=cfg.getboolean('LATTICE', 'variable_heading'),
        lon_straight_step=cfg.getfloat('LATTICE', 'lon_straight_step'),
        lon_curve_step=cfg.getfloat('LATTICE', 'lon_curve_step'),
        curve_thr=cfg.getfloat('LATTICE', 'curve_thr'),
        lat_offset=cfg.getfloat('LATTICE', 'lat_offset'),
        virt_goal_n=cfg.getboolean('LATTICE', 'virt_goal_n'),
        min_vel_race=cfg.getfloat('LATTICE', 'min_vel_race'),
        closure_detection_dist=cfg.getfloat('LATTICE', 'closure_detection_dist'),
        vel_decrease_lat=cfg.getfloat('PLANNINGTARGET', 'vel_decrease_lat'),
        min_plan_horizon=cfg.getfloat('PLANNINGTARGET', 'min_plan_horizon'),
        plan_horizon_mode=cfg.get('PLANNINGTARGET', 'plan_horizon_mode'),
        stepsize_approx=cfg.getfloat('SAMPLING', 'stepsize_approx'),
        veh_width=cfg.getfloat('VEHICLE', 'veh_width'),
        veh_length=cfg.getfloat('VEHICLE', 'veh_length'),
        veh_turn=cfg.getfloat('VEHICLE', 'veh_turn'),
        w_raceline=cfg.getfloat('COST', 'w_raceline'),
        w_raceline_sat=cfg.getfloat('COST', 'w_raceline_sat'),
        w_length=cfg.getfloat('COST', 'w_length'),
        w_curv_avg=cfg.getfloat('COST', 'w_curv_avg'),
        w_curv_peak=cfg.getfloat('COST', 'w_curv_peak'),
        w_virt_goal=cfg.getfloat('COST', 'w_virt_goal'))


def build_lattice(globtraj_input_path: str, offline_param_path: str = None, overrides: dict = None,
                  verbose: bool = False) -> Lattice:
    """Track CSV + offline ini (+ overrides of ini keys) -> Lattice.  Mirrors main_offline_callback.py:76-179."""
    if offline_param_path is not None:
        p = read_offline_config(offline_param_path)
        md5 = md5_of_files(globtraj_input_path, offline_param_path)
    else:
        p = {}
        md5 = md5_of_files(globtraj_input_path)
    if overrides:
        p.update(overrides)
        md5 += hashlib.md5(repr(sorted(overrides.items())).encode()).hexdigest()

    t = import_globtraj_csv(globtraj_input_path)
    refline, w_r, w_l, normvec, alpha = t["refline"], t["width_right"], t["width_left"], t["normvec"], t["alpha"]
    length_rl, vel_rl, kappa_rl = t["length_rl"], t["vel_rl"], t["kappa_rl"]

    # closed race line parameters: s, x, y, kappa, vel (main_offline_callback.py:89-104)
    s = np.concatenate(([0], np.cumsum(length_rl)))
    xy = refline + normvec * alpha[:, np.newaxis]
    rl_params = np.column_stack((xy, kappa_rl, vel_rl))
    closed = bool(np.hypot(xy[0, 0] - xy[-1, 0], xy[0, 1] - xy[-1, 1]) < p["closure_detection_dist"])
    if closed:
        glob_rl = np.column_stack((s, np.vstack((rl_params, rl_params[0, :]))))
    else:
        glob_rl = np.column_stack((s[:-1], rl_params))

    idx = variable_step_size(kappa=kappa_rl, dist=length_rl, d_curve=p["lon_curve_step"],
                             d_straight=p["lon_straight_step"], curve_th=p["curve_thr"], force_last=not closed)
    refline, w_r, w_l, normvec, alpha, vel_rl = refline[idx], w_r[idx], w_l[idx], normvec[idx], alpha[idx], vel_rl[idx]
    s_raceline = s[idx]
    num_layers = len(idx)

    g = _GraphParams()
    g.lat_offset = p["lat_offset"]
    g.lat_resolution = p["lat_resolution"]
    g.sampled_resolution = p["stepsize_approx"]
    g.vel_decrease_lat = p["vel_decrease_lat"]
    g.veh_width, g.veh_length, g.veh_turn = p["veh_width"], p["veh_length"], p["veh_turn"]
    g.virt_goal_node, g.virt_goal_node_cost = p["virt_goal_n"], p["w_virt_goal"]
    g.min_plan_horizon, g.plan_horizon_mode = p["min_plan_horizon"], p["plan_horizon_mode"]
    g.md5_params = md5
    g.s_raceline, g.vel_raceline, g.refline, g.normvec_normalized = s_raceline, vel_rl, refline, normvec
    g.track_width_right, g.track_width_left, g.alpha = w_r, w_l, alpha
    g.raceline = refline + normvec * alpha[:, np.newaxis]
    g.glob_rl = glob_rl

    if g.lat_offset <= 0:
        raise ValueError('Requested to small lateral offset! A lateral offset larger than zero must be allowed!')

    # ---- node skeleton (gen_node_skeleton.py:43-166; note: always evaluated with closed=True there) ----------------------
    raceline_pts = g.raceline
    psi = _heading_closed(raceline_pts)
    if p["variable_heading"]:
        bound_r = refline + normvec * w_r[:, None]
        bound_l = refline - normvec * w_l[:, None]
        psi_bl = _heading_closed(bound_l)
        psi_br = _heading_closed(bound_r)

    margin_left = np.min(w_l - g.veh_width / 2 + alpha)
    margin_right = np.min(w_r - g.veh_width / 2 - alpha)
    if margin_left < 0.0 or margin_right < 0.0:
        raise ValueError("Provided raceline holds points outside the safety margin! "
                         "Reduce the vehicle width or adapt the race line.")

    layer_pos, layer_psi, rl_index = [], [], []
    for i in range(num_layers):
        ri = int(np.floor((w_l[i] - g.veh_width / 2 + alpha[i]) / g.lat_resolution))
        rl_index.append(ri)
        s0 = alpha[i] - ri * g.lat_resolution
        alphas = np.arange(s0, w_r[i] - g.veh_width / 2, g.lat_resolution)
        pos = np.repeat(refline[i][None, :], len(alphas), axis=0) \
            + np.repeat(normvec[i][None, :], len(alphas), axis=0) * alphas[:, np.newaxis]
        if p["variable_heading"]:
            if abs(psi_bl[i] - psi[i]) < np.pi:
                psi1 = np.linspace(psi_bl[i], psi[i], num=ri + 1)[:-1]
            else:
                tb = psi_bl[i] + 2 * np.pi * (psi_bl[i] < 0)
                tp = psi[i] + 2 * np.pi * (psi[i] < 0)
                psi1 = _normalize_psi(np.linspace(tb, tp, num=ri + 1)[:-1])
            if abs(psi_br[i] - psi[i]) < np.pi:
                psi2 = np.linspace(psi[i], psi_br[i], num=len(alphas) - ri)
            else:
                tb = psi_br[i] + 2 * np.pi * (psi_br[i] < 0)
                tp = psi[i] + 2 * np.pi * (psi[i] < 0)
                psi2 = _normalize_psi(np.linspace(tp, tb, num=len(alphas) - ri))
            lpsi = np.append(psi1, psi2)
        else:
            lpsi = np.repeat(psi[i], len(alphas), axis=0)
        layer_pos.append(pos)
        layer_psi.append(lpsi)
    g.raceline_index = rl_index

    nil = np.array([lp.shape[0] for lp in layer_pos])
    node_off = np.concatenate(([0], np.cumsum(nil))).astype(np.int32)
    node_xy = np.concatenate(layer_pos, axis=0)
    node_psi = np.concatenate(layer_psi)

    # ---- edge candidates (gen_edges.py:52-105) -------------------------------------------------------------------------
    cx_r, cy_r = _closed_spline_coeffs(np.vstack((g.raceline, g.raceline[0])))
    e_sl, e_sn, e_en = [], [], []
    for sl in range(num_layers):
        el_ = sl + 1
        if el_ >= num_layers:
            if closed:
                el_ -= num_layers
            else:
                break
        n_end = nil[el_]
        for sn in range(nil[sl]):
            ref = rl_index[el_] + sn - rl_index[sl]
            d_start = layer_pos[sl][sn]
            d_end = layer_pos[el_][max(0, min(n_end - 1, ref))]
            dist = np.sqrt(np.power(d_end[0] - d_start[0], 2) + np.power(d_end[1] - d_start[1], 2))
            lat_steps = int(round(dist * g.lat_offset / g.lat_resolution))
            for en in range(max(0, ref - lat_steps), min(n_end, ref + lat_steps + 1)):
                e_sl.append(sl)
                e_sn.append(sn)
                e_en.append(en)
    e_sl = np.array(e_sl, dtype=np.int64)
    e_sn = np.array(e_sn, dtype=np.int64)
    e_en = np.array(e_en, dtype=np.int64)
    e_el = (e_sl + 1) % num_layers
    rl_arr = np.array(rl_index)
    is_rl = (rl_arr[e_sl] == e_sn) & (rl_arr[e_el] == e_en)
    gs = node_off[e_sl] + e_sn
    ge = node_off[e_el] + e_en
    cx, cy = _two_point_splines(node_xy[gs], node_xy[ge], node_psi[gs], node_psi[ge])
    cx[is_rl] = cx_r[e_sl[is_rl]]
    cy[is_rl] = cy_r[e_sl[is_rl]]
    if verbose:
        print("lattice: %d layers, %d nodes, %d edge candidates" % (num_layers, node_off[-1], e_sl.size))

    # ---- sampling (gen_edges.py:113-156; tph.interp_splines(stepsize_approx) + calc_head_curv_an) ------------------------
    t15 = np.linspace(0.0, 1.0, 15)
    px = cx[:, 0:1] + cx[:, 1:2] * t15 + cx[:, 2:3] * np.power(t15, 2) + cx[:, 3:4] * np.power(t15, 3)
    py = cy[:, 0:1] + cy[:, 1:2] * t15 + cy[:, 2:3] * np.power(t15, 2) + cy[:, 3:4] * np.power(t15, 3)
    seg = np.sqrt(np.power(np.diff(px, axis=1), 2) + np.power(np.diff(py, axis=1), 2))
    spl_len = np.array([np.sum(row) for row in seg])
    npts = np.ceil(spl_len / g.sampled_resolution).astype(np.int64) + 1
    samp_off = np.concatenate(([0], np.cumsum(npts)))
    eidx = np.repeat(np.arange(e_sl.size), npts)
    k_in = np.arange(samp_off[-1]) - samp_off[eidx]
    # dists_interp = linspace(0, len, n): start + k * step with step = len / (n - 1); t = dist / len; last point t = 1
    step = spl_len / (npts - 1)
    tt = (k_in * step[eidx]) / spl_len[eidx]
    last = k_in == (npts[eidx] - 1)
    tt[last] = 1.0
    a = cx[eidx]
    b = cy[eidx]
    sx = a[:, 0] + a[:, 1] * tt + a[:, 2] * np.power(tt, 2) + a[:, 3] * np.power(tt, 3)
    sy = b[:, 0] + b[:, 1] * tt + b[:, 2] * np.power(tt, 2) + b[:, 3] * np.power(tt, 3)
    sx[last] = np.sum(cx, axis=1)[eidx[last]]
    sy[last] = np.sum(cy, axis=1)[eidx[last]]
    xd = a[:, 1] + 2 * a[:, 2] * tt + 3 * a[:, 3] * np.power(tt, 2)
    yd = b[:, 1] + 2 * b[:, 2] * tt + 3 * b[:, 3] * np.power(tt, 2)
    xdd = 2 * a[:, 2] + 6 * a[:, 3] * tt
    ydd = 2 * b[:, 2] + 6 * b[:, 3] * tt
    spsi = _normalize_psi(np.arctan2(yd, xd) - math.pi / 2)
    skap = (xd * ydd - yd * xdd) / np.power(np.power(xd, 2) + np.power(yd, 2), 1.5)

    # ---- turn radius / race speed filter (gen_edges.py:138-156) -----------------------------------------------------------
    vel_min = vel_rl[e_sl] * p["min_vel_race"]
    with np.errstate(divide="ignore"):
        kap_lim = np.minimum(1.0 / g.veh_turn, 1.0 / (np.power(vel_min, 2) / 10.0))
    viol = np.abs(skap) > kap_lim[eidx]
    viol_edge = np.bincount(eidx[viol], minlength=e_sl.size) > 0
    keep = (~viol_edge) | is_rl

    # ---- prune dead ends (prune_graph.py:24-67): greatest sub-graph where every node with an edge has in- and out-edges ----
    nn = int(node_off[-1])
    node_layer = np.repeat(np.arange(num_layers), nil)
    protected = np.zeros(nn, dtype=bool)
    if not closed:
        protected = (node_layer == 0) | (node_layer == num_layers - 1)
    while True:
        outdeg = np.bincount(gs[keep], minlength=nn)
        indeg = np.bincount(ge[keep], minlength=nn)
        dead = ((outdeg == 0) | (indeg == 0)) & ~protected
        rm = keep & (dead[gs] | dead[ge])
        if not rm.any():
            break
        keep &= ~rm

    # ---- per-edge records + offline cost (GB:425-436, gen_offline_cost.py:47-79) -----------------------------------------
    recs = []
    kept = np.nonzero(keep)[0]
    for e in kept:
        a0, a1 = samp_off[e], samp_off[e + 1]
        xyk = np.column_stack((sx[a0:a1], sy[a0:a1], spsi[a0:a1], skap[a0:a1]))
        el = np.sqrt(np.sum(np.power(np.diff(xyk[:, 0:2], axis=0), 2), axis=1))
        slen = np.sum(el)
        param = np.column_stack((xyk, np.append(el, 0)))
        kap = param[:, 3]
        cost = 0.0
        cost += p["w_curv_avg"] * np.power(sum(abs(kap)) / float(len(kap)), 2) * slen
        cost += p["w_curv_peak"] * np.power(abs(max(kap) - min(kap)), 2) * slen
        cost += p["w_length"] * slen
        rdist = abs(rl_index[int(e_el[e])] - int(e_en[e])) * g.lat_resolution
        cost += min(p["w_raceline"] * slen * rdist, p["w_raceline_sat"] * slen)
        recs.append((int(e_sl[e]), int(e_en[e]), int(e_sn[e]), float(cost), float(slen), param))

    return _assemble(num_layers=num_layers, closed=closed, node_off=node_off, node_x=node_xy[:, 0].copy(),
                     node_y=node_xy[:, 1].copy(), node_psi=node_psi, recs=recs, gb_like=g)


def load_or_build_lattice(globtraj_input_path: str, offline_param_path: str, store_path: str = None,
                          overrides: dict = None, force_recalc: bool = False) -> tuple:
    """md5-keyed cache like main_offline_callback.py:57-68,183-185 (flat .npz instead of an igraph pickle)."""
    md5 = md5_of_files(globtraj_input_path, offline_param_path)
    if overrides:
        md5 += hashlib.md5(repr(sorted(overrides.items())).encode()).hexdigest()
    if store_path is not None and not force_recalc and os.path.isfile(store_path):
        try:
            lat = Lattice.load(store_path)
            if lat.md5_params == md5:
                return lat, False
        except Exception:
            pass
    lat = build_lattice(globtraj_input_path, offline_param_path, overrides=overrides)
    if store_path is not None:
        try:
            lat.save(store_path)
        except OSError:
            pass
    return lat, True
