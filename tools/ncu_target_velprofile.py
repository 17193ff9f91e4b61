"""ncu target (GPU box): BASELINE config 5, 100 k paths x 500 points through ltpl_velprofile_batch.
ncu --set full --clock-control none --import-source on -k regex:k_velprofile --launch-skip 2 --launch-count 1
    -o gpurun_out/prof_vp python tools/ncu_target_velprofile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
from graphbasedlocaltrajectoryplanner_b200.scenarios import make_velocity_microbench
from graphbasedlocaltrajectoryplanner_b200.velprofile import velprofile_batch_device
pl = BatchPlanner(bench.get_lattice("l216"), device="cuda:0")
pl.set_vel_params(vel_max=60.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=bench.ax_max_machines(), safety_d=30.0)
mb = make_velocity_microbench(100000, 500)
d = {k: torch.from_numpy(np.ascontiguousarray(mb[k])).cuda() for k in ("kappa", "el", "v_start", "v_end")}
vx = torch.empty_like(d["kappa"]); ax = torch.empty_like(d["kappa"])
for _ in range(4):
    velprofile_batch_device(pl, d["kappa"], d["el"], d["v_start"], d["v_end"], vx, ax)
torch.cuda.synchronize()
