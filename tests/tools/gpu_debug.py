"""debug helper (GPU box): per-scenario diff of the CUDA path against the oracle with details."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import helpers as H
from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
from graphbasedlocaltrajectoryplanner_b200.scenarios import Track, make_scenarios
from graphbasedlocaltrajectoryplanner_b200 import capi
from oracle.ltpl_oracle import OracleLTPL

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 384
g = H.golden("ticks_%s.npz" % tag)
axm = g["ax_max_machines"]
VEL = dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), safety_d=30.0)
sc = make_scenarios(Track(H.TRACK_CSV), n, seed=4242 + n, n_obj_min=0 if tag == "default" else 1, n_obj_max=3)
pl = BatchPlanner(H.lattice_for(tag), device="cuda:0")
pl.set_vel_params(ax_max_machines=axm, **VEL)
pl.stage_scenarios(sc); pl.upload(); pl.set_startpos(); pl.calc_paths(); pl.calc_vel_profile()
recs = pl.records()
f = pl.fetch("s_vx_ax", "status", "action_id", "path_len")
orc = OracleLTPL(H.lattice_for(tag))
vk = dict(ax_max_machines=axm, **VEL)
shown = 0
for b in range(sc.size):
    want = orc.tick(sc.pos[b], sc.heading[b], sc.vel[b], sc.object_list(b), vk)
    try:
        H.compare_records(recs[b], want, ctx="%s %d" % (tag, b))
    except AssertionError as e:
        shown += 1
        print("FAIL", str(e).split("\n")[0][:200])
        print("  vel", sc.vel[b], "oracle vel_bound", want.get("vel_bound"))
        for s in range(3):
            a = int(f["action_id"][s, b]); q = s * sc.size + b
            if a < 0: continue
            npts = int(f["path_len"][s, b])
            print("  slot", s, capi.ACTION_NAMES[a], "status", bin(int(f["status"][s, b])), "n", npts,
                  "vx[:6]", np.round(f["s_vx_ax"][1, q, :6], 4), "vx[-3:]", np.round(f["s_vx_ax"][1, q, npts-3:npts], 4))
        if shown >= 6: break
print("done; failures shown:", shown)
