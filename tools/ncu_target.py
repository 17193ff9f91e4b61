"""ncu target (GPU box): set_startpos + 3 ticks of the bench workload.
ncu --set full --clock-control none --import-source on -k regex:'k_plan|k_path|k_vel_res|k_export' --launch-skip 8
    --launch-count 4 -o gpurun_out/prof python tools/ncu_target.py l216"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
tag = sys.argv[1] if len(sys.argv) > 1 else "l216"
pl = BatchPlanner(bench.get_lattice(tag), device="cuda:0")
pl.set_vel_params(**bench.vel_kwargs())
pl.stage_scenarios(bench.make_batch(tag, 10000)); pl.upload(); pl.set_startpos()
for _ in range(3):
    pl.tick()
torch.cuda.synchronize()
