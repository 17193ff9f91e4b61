"""
Multi-GPU plumbing (one process per GPU, torch.distributed).

The planning path shards trivially: scenarios are independent, there is NO data-path collective.  The two
collectives are bookkeeping around it:
  * ``broadcast_lattice``  -- rank 0 packs the read-only lattice blob once, every other rank receives the bytes
    (NCCL broadcast on the device, over NVLink 5 / NVSwitch) and builds its own handle on them;
  * ``gather_action_sets`` -- fixed-stride exported trajectories of every rank's shard are all-gathered.
Both work with the ``gloo`` backend on CPU tensors too (world_size-2 tests without a GPU).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .lattice_blob import pack_lattice

_CAP_KEYS = ("h_max", "max_window_edges", "p0_max", "p_max", "max_plan_layers")


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """scenario i is planned by rank i % world."""
    return np.arange(rank, n, world)


def broadcast_lattice(lattice, device, src: int = 0) -> tuple:
    """returns (LatticeHeader, capacities dict, blob tensor on `device`).  `lattice` is only needed on rank `src`."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    hsize = C.sizeof(capi.LatticeHeader)
    if rank == src:
        header, blob, cap = pack_lattice(lattice)
        meta = np.zeros(hsize + 8 * len(_CAP_KEYS), dtype=np.uint8)
        meta[:hsize] = np.frombuffer(bytes(header), dtype=np.uint8)
        meta[hsize:] = np.array([cap[k] for k in _CAP_KEYS], dtype=np.int64).view(np.uint8)
        meta_t = torch.from_numpy(meta).to(device)
    else:
        meta_t = torch.zeros(hsize + 8 * len(_CAP_KEYS), dtype=torch.uint8, device=device)
    if world > 1:
        dist.broadcast(meta_t, src=src)
    meta_h = meta_t.cpu().numpy()
    header = capi.LatticeHeader.from_buffer_copy(meta_h[:hsize].tobytes())
    cap = dict(zip(_CAP_KEYS, (int(v) for v in meta_h[hsize:].view(np.int64))))
    if rank == src:
        blob_t = torch.from_numpy(blob).to(device)
    else:
        blob_t = torch.empty(int(header.blob_bytes), dtype=torch.uint8, device=device)
    if world > 1:
        dist.broadcast(blob_t, src=src)
    return header, cap, blob_t


def gather_action_sets(traj: torch.Tensor, traj_len: torch.Tensor, traj_id: torch.Tensor, out: list = None) -> tuple:
    """all-gather the fixed-stride exported trajectories ([rows][n_export][7] fp32 + lengths + ids) on the current
    stream.  Every rank must hold the same local sizes.  Returns tensors with a leading world dimension; ``out`` (a list,
    filled on first use) keeps the receive buffers across calls."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return traj.unsqueeze(0), traj_len.unsqueeze(0), traj_id.unsqueeze(0)
    world = dist.get_world_size()
    outs = []
    for i, t in enumerate((traj, traj_len, traj_id)):
        shape = (world * t.shape[0],) + tuple(t.shape[1:])
        if out is not None and len(out) > i and tuple(out[i].shape) == shape and out[i].dtype == t.dtype:
            buf = out[i]
        else:
            buf = torch.empty(shape, dtype=t.dtype, device=t.device)
            if out is not None:
                if len(out) > i:
                    out[i] = buf
                else:
                    out.append(buf)
        dist.all_gather_into_tensor(buf, t.contiguous())   # concatenated along dim 0 (valid for nccl and gloo)
        outs.append(buf.view((world,) + tuple(t.shape)))
    return tuple(outs)
