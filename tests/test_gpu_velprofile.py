"""stand-alone forward/backward velocity solver (BASELINE config 5 shape, small) against the tph restatement."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_velprofile_dense_matches_tph_port():
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
    from graphbasedlocaltrajectoryplanner_b200.scenarios import make_velocity_microbench
    from graphbasedlocaltrajectoryplanner_b200.velprofile import calc_vel_profile_batch
    from oracle import tph_port as tph
    g = H.golden("ticks_default.npz")
    axm = g["ax_max_machines"]
    mb = make_velocity_microbench(256, 500, seed=3)
    pl = BatchPlanner(H.lattice_for("l216"), device="cuda:0")
    pl.set_vel_params(vel_max=60.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=axm, safety_d=30.0)
    vx, ax = calc_vel_profile_batch(pl, mb["kappa"], mb["el"], mb["v_start"], mb["v_end"])
    worst_v, worst_a = 0.0, 0.0
    for i in range(256):
        ref = tph.calc_vel_profile(ax_max_machines=axm, kappa=mb["kappa"][i], el_lengths=mb["el"][i, :-1], closed=False,
                                   drag_coeff=0.85, m_veh=1000.0, loc_gg=np.ones((500, 2)) * 5.0, v_max=60.0,
                                   v_start=mb["v_start"][i], v_end=mb["v_end"][i])
        ref_ax = tph.calc_ax_profile(ref, mb["el"][i, :-1])
        worst_v = max(worst_v, float(np.max(np.abs(vx[i] - ref) / (1e-3 + np.abs(ref)))))
        worst_a = max(worst_a, float(np.max(np.abs(ax[i, :-1] - ref_ax))))
    assert worst_v < 1e-4, worst_v       # 1e-4 relative (north_star)
    assert worst_a < 5e-3, worst_a


def test_velprofile_ragged_sizes_and_general_exponent():
    """70 paths x 137 points (neither a multiple of the warp nor of the tile), friction-ellipse exponent 1.5, other mass /
    drag: the tile edges and the pow path of the fp32 recurrences against the tph restatement."""
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
    from graphbasedlocaltrajectoryplanner_b200.scenarios import make_velocity_microbench
    from graphbasedlocaltrajectoryplanner_b200.velprofile import calc_vel_profile_batch
    from oracle import tph_port as tph
    axm = H.golden("ticks_default.npz")["ax_max_machines"]
    mb = make_velocity_microbench(70, 137, seed=4)
    pl = BatchPlanner(H.lattice_for("l216"), device="cuda:0", veh_param_dyn_model_exp=1.5, veh_param_dragcoeff=0.9,
                      veh_param_mass=1200.0)
    pl.set_vel_params(vel_max=55.0, gg_scale=0.9, local_gg=(4.5, 5.5), ax_max_machines=axm, safety_d=30.0)
    vx, ax = calc_vel_profile_batch(pl, mb["kappa"], mb["el"], mb["v_start"], mb["v_end"])
    for i in range(70):
        ref = tph.calc_vel_profile(ax_max_machines=axm, kappa=mb["kappa"][i], el_lengths=mb["el"][i, :-1], closed=False,
                                   drag_coeff=0.9, m_veh=1200.0, loc_gg=np.ones((137, 2)) * (4.5, 5.5) * 0.9, v_max=55.0,
                                   v_start=mb["v_start"][i], v_end=mb["v_end"][i], dyn_model_exp=1.5)
        ref_ax = tph.calc_ax_profile(ref, mb["el"][i, :-1])
        assert np.max(np.abs(vx[i] - ref) / (1e-3 + np.abs(ref))) < 1e-4, "path %d vx" % i
        assert np.max(np.abs(ax[i, :-1] - ref_ax)) < 5e-3, "path %d ax" % i
        assert ax[i, -1] == 0.0
