"""BASELINE.json's full batch size on the GPU -- 10 000 scenarios on the ~200 x 11 lattice with 1-3 objects (configs[1],
the bench workload) and on the 400 x 21 lattice with 5 objects (configs[3]) -- through properties that do not need the
oracle for every scenario:
  * a scenario's result does not depend on the batch it is planned in: any sub-batch and any permutation of the batch
    reproduce the same per-scenario bytes (this is also what lets the batch shard across GPUs without a collective);
  * multiset of per-scenario checksums of the 4 rank shards (ScenarioBatch.shard, the N-GPU split) == that of the full batch;
  * structural invariants of every returned node sequence and trajectory (layer succession, arc length, velocity bounds,
    ax consistent with vx, acceptance rule OTH:907-911, export cut LTPL:401-406);
  * the oracle as the checker on a random sample of the big batch."""
import zlib

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

VEL = dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), safety_d=30.0)
B_FULL = 10000


def _plan(tag, sc, axm):
    """fresh planner (same tick counter => same trajectory ids), one first tick; canonical per-scenario arrays."""
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
    pl = BatchPlanner(H.lattice_for(tag), device="cuda:0")
    pl.set_vel_params(ax_max_machines=axm, **VEL)
    pl.stage_scenarios(sc)
    pl.upload()
    pl.set_startpos()
    pl.tick()
    f = pl.fetch("sc_flags", "start_node", "action_id", "status", "n_nodes", "nodes", "path_len", "traj_len", "traj_id",
                 "traj_row", "traj", "closest_obj")
    ns, B = f["action_id"].shape
    ne = f["traj"].shape[1]
    nodes = f["nodes"].copy()
    nodes[np.arange(nodes.shape[2])[None, None, :] >= f["n_nodes"][..., None]] = -1     # stale entries behind the list
    traj = np.zeros((ns, B, ne, 7), dtype=np.float32)
    ok = f["traj_row"] >= 0
    traj[ok] = f["traj"][f["traj_row"][ok]]
    traj[np.arange(ne)[None, None, :] >= f["traj_len"][..., None]] = 0.0               # rows behind the export cut
    can = dict(flags=f["sc_flags"], start_node=f["start_node"], closest_obj=f["closest_obj"],
               action_id=f["action_id"].T, status=f["status"].T, n_nodes=f["n_nodes"].T, path_len=f["path_len"].T,
               traj_len=f["traj_len"].T, traj_id=f["traj_id"].T, nodes=nodes.transpose(1, 0, 2, 3),
               traj=traj.transpose(1, 0, 2, 3))
    return pl, {k: np.ascontiguousarray(v) for k, v in can.items()}


def _checksums(can):
    keys = sorted(can)
    B = can["flags"].shape[0]
    return np.array([zlib.crc32(b"".join(can[k][b].tobytes() for k in keys)) for b in range(B)], dtype=np.uint64)


def _assert_same(a, b, idx, what):
    for k in sorted(a):
        same = np.array_equal(a[k][idx], b[k])
        if not same:
            rows = np.nonzero([not np.array_equal(x, y) for x, y in zip(a[k][idx], b[k])])[0]
            raise AssertionError("%s: '%s' differs for %d scenarios (first: batch index %d)" % (
                what, k, rows.size, int(np.asarray(idx)[rows[0]])))


@pytest.mark.parametrize("tag,omin,omax,n_oracle", [("l216", 1, 3, 48), ("l430", 5, 5, 16)])
def test_full_batch_properties(tag, omin, omax, n_oracle):
    from graphbasedlocaltrajectoryplanner_b200 import capi
    from graphbasedlocaltrajectoryplanner_b200.scenarios import Track, make_scenarios
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("ticks_%s.npz" % tag)
    axm = g["ax_max_machines"]
    lat = H.lattice_for(tag)
    sc = make_scenarios(Track(H.track_csv_for(tag)), B_FULL, seed=777, n_obj_min=omin, n_obj_max=omax)
    pl, full = _plan(tag, sc, axm)
    rng = np.random.default_rng(778)

    # ---- batch composition does not matter -------------------------------------------------------------------------------
    idx = np.sort(rng.choice(B_FULL, size=777, replace=False))          # ragged size on purpose
    _assert_same(full, _plan(tag, sc.subset(idx), axm)[1], idx, "sub-batch of 777")
    perm = rng.permutation(B_FULL)
    _assert_same(full, _plan(tag, sc.subset(perm), axm)[1], perm, "permuted batch")
    sums = _checksums(full)
    shard_sums = np.concatenate([_checksums(_plan(tag, sc.shard(r, 4), axm)[1]) for r in range(4)])
    assert np.array_equal(np.sort(sums), np.sort(shard_sums)), "checksums of the 4 rank shards != full batch"

    # ---- structural invariants of every result ---------------------------------------------------------------------------
    okb = full["flags"] == 0
    assert okb.mean() > 0.99, "flagged scenarios: %d" % int((~okb).sum())
    act, st = full["action_id"], full["status"]
    assert np.all(act[~okb] == capi.ACT_NONE) and np.all(full["traj_len"][~okb] == 0)
    has = act != capi.ACT_NONE
    assert has[okb].any(axis=1).mean() > 0.99
    # slot 0 holds follow or straight; left / right never come without follow (MOPG:124-174)
    assert np.all(np.isin(act[:, 0][has[:, 0]], (capi.ACT_STRAIGHT, capi.ACT_FOLLOW)))
    lr = np.isin(act, (capi.ACT_LEFT, capi.ACT_RIGHT)).any(axis=1)
    assert np.all(act[lr, 0] == capi.ACT_FOLLOW)
    # node sequences: behind the first entry every node lies one layer further (closed track: modulo)
    L = int(lat.num_layers)
    nn, nodes = full["n_nodes"], full["nodes"]
    found = has & ((st & capi.ST_FOUND) != 0)
    assert np.all(nn[found] >= 2)
    lay = nodes[..., 0].astype(np.int64)
    i = np.arange(lay.shape[2] - 1)[None, None, :]
    chk = found[..., None] & (i >= 1) & (i + 1 < nn[..., None])
    assert np.all(((lay[..., :-1] + 1) % L == lay[..., 1:])[chk]), "layer succession broken"
    assert np.all(nodes[..., 1][(np.arange(lay.shape[2])[None, None, :] < nn[..., None]) & found[..., None]
                                & (np.arange(lay.shape[2])[None, None, :] >= 1)] >= 0)
    # trajectories
    tl, tr = full["traj_len"], full["traj"].astype(np.float64)
    ne = tr.shape[2]
    assert tl.max() <= ne and np.all(tl[~has] == 0)
    valid = tl > 0
    assert np.all((st[valid] & capi.ST_TRAJ_VALID) != 0) and valid.sum() > B_FULL
    j = np.arange(ne)[None, None, :]
    inside = j < tl[..., None]
    s, vx, ax = tr[..., 0], tr[..., 5], tr[..., 6]
    assert np.all(s[..., 0][valid] == 0.0)
    step = inside[..., 1:]
    ds = np.diff(s, axis=2)
    assert np.all(ds[step] > 0.0), "arc length not increasing"
    assert np.all(vx[inside] >= 0.0) and np.all(vx[inside] <= VEL["vel_max"] + 1e-3)
    # ax = (v1^2 - v0^2) / (2 ds) (OTH:935-936), -5 where the vehicle stands (OTH:939); fp32 export
    ax_want = (vx[..., 1:] ** 2 - vx[..., :-1] ** 2) / (2.0 * np.where(step, ds, 1.0))
    moving = step & ~((np.abs(vx[..., :-1]) < 1e-6) & (np.abs(ax_want) < 1e-6))
    err = np.abs(ax[..., :-1] - ax_want)[moving]
    lim = 2e-2 + 2e-3 * np.abs(ax_want[moving]) + 1e-2 / ds[moving]      # fp32 rounding of vx^2 and s in the export
    assert np.all(err <= lim), "ax inconsistent with vx: %d rows, worst %.3e" % (int((err > lim).sum()), err.max())
    # acceptance (OTH:907-911): a kept left / right / straight profile starts at the planned velocity (follow mode has its
    # own flag, CVPF:78-313); follow / straight are kept regardless and carry the violation bit
    v0_ok = np.abs(vx[..., 0] - sc.vel[:, None]) < 0.1 + 1e-3
    viol = (st & capi.ST_VEL_BOUND_VIOL) != 0
    assert np.all(v0_ok[valid & ~viol & (act != capi.ACT_FOLLOW)])
    assert not np.any(viol & valid & np.isin(act, (capi.ACT_LEFT, capi.ACT_RIGHT)))
    # ids: base (10 * tick) + action offset, distinct within a scenario
    ids = full["traj_id"]
    assert np.all(ids[valid] // 10 == 1)
    for a in range(ids.shape[1]):
        for c in range(a + 1, ids.shape[1]):
            both = valid[:, a] & valid[:, c]
            assert np.all(ids[both, a] != ids[both, c])

    # ---- the oracle on a sample of the big batch -------------------------------------------------------------------------
    pick = np.sort(rng.choice(B_FULL, size=n_oracle, replace=False))
    recs = pl.records(indices=pick.tolist())
    orc = OracleLTPL(lat)
    vk = dict(ax_max_machines=axm, **VEL)
    fails = []
    for rec, b in zip(recs, pick):
        try:
            H.compare_records(rec, orc.tick(sc.pos[b], sc.heading[b], sc.vel[b], sc.object_list(int(b)), vk),
                              ctx="%s scenario %d of %d" % (tag, b, B_FULL))
        except AssertionError as e:
            fails.append(str(e).split("\n")[0][:300])
    assert not fails, "%d/%d sampled scenarios differ from the oracle:\n%s" % (len(fails), n_oracle, "\n".join(fails[:8]))
