"""compute-sanitizer target (GPU box): one small tick of every kernel on each lattice + the dense velocity microbench.
Run as  compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/gpu_sanitize.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import helpers as H
from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
from graphbasedlocaltrajectoryplanner_b200.scenarios import Track, make_scenarios, make_velocity_microbench
from graphbasedlocaltrajectoryplanner_b200.velprofile import calc_vel_profile_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
VEL = dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), safety_d=30.0)
for tag in ("default", "l216"):
    g = H.golden("ticks_%s.npz" % tag)
    sc = make_scenarios(Track(H.TRACK_CSV), n, seed=77, n_obj_min=0, n_obj_max=3)
    pl = BatchPlanner(H.lattice_for(tag), device="cuda:0")
    pl.set_vel_params(ax_max_machines=g["ax_max_machines"], **VEL)
    pl.stage_scenarios(sc); pl.upload(); pl.set_startpos(); pl.tick()
    r = pl.records()
    print(tag, "trajectories:", sum(len(x.get("traj", {})) for x in r))
mb = make_velocity_microbench(200, 150, seed=3)
vx, ax = calc_vel_profile_batch(pl, mb["kappa"], mb["el"], mb["v_start"], mb["v_end"])
print("dense vx mean", float(np.mean(vx)))
