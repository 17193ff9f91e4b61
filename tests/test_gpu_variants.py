"""parameter variants of the velocity planner / follow controller through the C-ABI against golden vectors of the
unmodified reference (tests/golden/ticks_variants_default.npz): PDtan controller, friction-ellipse exponents 1.5 / 2.0
(pow path of the tyre model), other mass / drag, gg scale, asymmetric gg, ego velocity estimate != planned velocity."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(H.VARIANTS))
def test_parameter_variant_matches_reference_golden(name):
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
    from graphbasedlocaltrajectoryplanner_b200.scenarios import ScenarioBatch
    g = H.golden("ticks_variants_default.npz")
    sub = H._Sub(g, name)
    online, veh, vel, dv = H.VARIANTS[name]
    pl = BatchPlanner(H.lattice_for("default"), online=online, device="cuda:0", **veh)
    pl.set_vel_params(ax_max_machines=g["ax_max_machines"], **vel)
    n = sub["sc_pos"].shape[0]
    sc = ScenarioBatch.from_object_lists(sub["sc_pos"], sub["sc_heading"], sub["sc_vel"],
                                         [H.object_list(sub, b) for b in range(n)], k_max=3)
    pl.stage_scenarios(sc, vel_est=sub["sc_vel"] + dv)
    pl.upload()
    pl.set_startpos()
    pl.tick()
    recs = pl.records()
    n_follow = 0
    for b in range(n):
        H.compare_record(recs[b], sub, b, prefix="", ctx=name + " gpu")
        n_follow += int("follow" in recs[b].get("traj", {}))
    assert n_follow >= n // 2
