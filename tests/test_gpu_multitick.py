"""EXPERIMENTAL stateful tick on the device (csrc/ltpl_state.cuh, BatchPlanner.next_tick) against the closed-loop
sequences of the unmodified reference (tests/golden/ticks_multitick_default.npz, scripted clock): the 16 sequences run
as ONE batch, tick 0 = set_startpos + first tick, ticks 1.. = next_tick with the recorded inputs."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _t_const(dts):
    """OTH:353-375: moving average (5) of the calculation times * calc_time_safety (2.0), capped at 0.5 s."""
    buf, out = [], []
    for dt in dts:
        if len(buf) >= 5:
            buf.pop(0)
        buf.append(float(dt))
        out.append(min(float(np.sum(buf) / len(buf)) * 2.0, 0.5))
    return out


def test_next_tick_matches_reference_sequences():
    from graphbasedlocaltrajectoryplanner_b200 import capi
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
    from graphbasedlocaltrajectoryplanner_b200.scenarios import ScenarioBatch
    g = H.golden("ticks_multitick_default.npz")
    n_seq, n_ticks = g["dt"].shape
    assert int(g["n_done"].min()) == n_ticks
    pl = BatchPlanner(H.lattice_for("default"), device="cuda:0", stateful=True)
    pl.set_vel_params(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=g["ax_max_machines"],
                      safety_d=30.0)
    tc = np.array([_t_const(g["dt"][q, 1:]) for q in range(n_seq)])       # t_const of ticks 1 ..
    fails, compared = [], 0
    alive = np.ones(n_seq, dtype=bool)
    for k in range(n_ticks):
        sc = ScenarioBatch(g["pos_est"][:, k].copy(), g["sc_heading"].copy(), g["sc_vel"].copy(), g["sc_n_obj"].copy(),
                           g["obj"][:, k].copy())
        if k == 0:
            pl.stage_scenarios(sc, vel_est=g["vel_est"][:, k])
            pl.upload()
            pl.set_startpos()
            pl.tick()
        else:
            pl.next_tick(sc, sel_action=g["sel"][:, k], t_const=tc[:, k - 1], vel_est=g["vel_est"][:, k])
        recs = pl.records()
        for q in range(n_seq):
            if not alive[q]:
                continue
            ctx = "sequence %d tick %d" % (q, k)
            rec = recs[q]
            try:
                assert not (rec["flags"] & capi.SC_STATE_FALLBACK), ctx + " fell back (flags %d)" % rec["flags"]
                assert not rec["out_of_track"] and "error" not in rec, ctx + " flags %d" % rec["flags"]
                for a, act in enumerate(H.ACTIONS):
                    n_want = int(g["path_len"][q, k, a])
                    has = act in rec["paths"]
                    assert has == (n_want > 0), "%s: path %s present=%s, golden %d" % (ctx, act, has, n_want)
                    if has and not rec["tie"].get(act):
                        nd = [[-1 if v is None else int(v) for v in p] for p in rec["nodes"][act][0]]
                        want = g["nodes"][q, k, a, :int(g["nodes_len"][q, k, a])].tolist()
                        assert nd == want, "%s: nodes of %s\\n got  %s\\n want %s" % (ctx, act, nd, want)
                        assert rec["paths"][act][0].shape[0] == n_want, "%s: path length %s %d vs %d" % (
                            ctx, act, rec["paths"][act][0].shape[0], n_want)
                    t_want = int(g["traj_len"][q, k, a])
                    t_has = act in rec["traj"]
                    assert t_has == (t_want > 0), "%s: trajectory %s present=%s, golden %d" % (ctx, act, t_has, t_want)
                    if t_has:
                        assert rec["traj"][act][0].shape[0] == t_want, "%s: rows of %s %d vs %d" % (
                            ctx, act, rec["traj"][act][0].shape[0], t_want)
                        H.assert_close("traj[%s]" % act, rec["traj"][act][0], g["traj"][q, k, a, :t_want],
                                       ("s", "x", "y", "psi", "kappa", "vx", "ax"), ctx)
                        compared += 1
            except AssertionError as e:
                fails.append(str(e).split("\\n")[0][:400] if "nodes of" not in str(e) else str(e)[:700])
                alive[q] = False            # later ticks of this sequence depend on this one
    assert not fails, "%d sequences diverged (of %d; %d trajectories matched before):\\n%s" % (
        len(fails), n_seq, compared, "\\n".join(fails[:8]))
    assert compared > 150
