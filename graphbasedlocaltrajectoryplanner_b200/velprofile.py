"""
Stand-alone batched forward/backward ggv velocity profile (BASELINE.json config 5).

Replaces ``VpForwardBackward.calc_vel_profile`` -> ``tph.calc_vel_profile(closed=False)`` with ``loc_gg`` constant
along the path (/root/reference/graph_ltpl/online_graph/src/VpForwardBackward.py:194-227) for N dense paths at once,
through ``ltpl_velprofile_batch`` of the C-ABI.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi


def velprofile_batch_device(planner, kappa: torch.Tensor, el: torch.Tensor, v_start: torch.Tensor, v_end: torch.Tensor,
                            vx: torch.Tensor, ax: torch.Tensor) -> None:
    """device tensors in / out (float64, [n_paths][n_points], contiguous); vehicle / gg parameters from
    ``planner.params`` (set_vel_params)."""
    for t in (kappa, el, vx, ax):
        if t.dtype != torch.float64 or not t.is_contiguous() or t.shape != kappa.shape:
            raise ValueError("kappa, el, vx, ax must be contiguous float64 tensors of identical shape")
    vb = capi.VelBatch()
    vb.n_paths, vb.n_points = int(kappa.shape[0]), int(kappa.shape[1])
    vb.kappa, vb.el, vb.v_start, vb.v_end = kappa.data_ptr(), el.data_ptr(), v_start.data_ptr(), v_end.data_ptr()
    vb.vx, vb.ax = vx.data_ptr(), ax.data_ptr()
    capi.check(planner.lib, planner.lib.ltpl_velprofile_batch(C.byref(planner.params), C.byref(vb), planner.stream),
               "ltpl_velprofile_batch")


def calc_vel_profile_batch(planner, kappa: np.ndarray, el_lengths: np.ndarray, v_start: np.ndarray,
                           v_end: np.ndarray) -> tuple:
    """host arrays in, host arrays out: kappa [N][P], el_lengths [N][P-1] or [N][P]; returns (vx [N][P], ax [N][P])."""
    kappa = np.ascontiguousarray(kappa, dtype=np.float64)
    n, p = kappa.shape
    el = np.zeros((n, p))
    el[:, :el_lengths.shape[1]] = el_lengths
    dev = planner.device
    d_k = torch.from_numpy(kappa).to(dev)
    d_e = torch.from_numpy(el).to(dev)
    d_s = torch.from_numpy(np.ascontiguousarray(v_start, dtype=np.float64)).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(v_end, dtype=np.float64)).to(dev)
    vx = torch.empty_like(d_k)
    ax = torch.empty_like(d_k)
    velprofile_batch_device(planner, d_k, d_e, d_s, d_t, vx, ax)
    torch.cuda.synchronize(dev)
    return vx.cpu().numpy(), ax.cpu().numpy()
