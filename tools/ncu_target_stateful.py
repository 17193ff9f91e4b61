"""ncu target (GPU box): the bench workload as a closed loop -- first tick, then 3 stateful ticks (next_tick), every
scenario driving 0.1 s along its first kept trajectory (the loop of bench.py extra.stateful_tick).
ncu --set full --clock-control none -k regex:'k_state|k_plan|k_path|k_ref|k_vel_res|k_backup|k_prefix|k_export'
    --launch-skip 14 --launch-count 8 -o gpurun_out/prof_st python tools/ncu_target_stateful.py l216"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
from graphbasedlocaltrajectoryplanner_b200.scenarios import ScenarioBatch
tag = sys.argv[1] if len(sys.argv) > 1 else "l216"
dt = 0.1
sc = bench.make_batch(tag, 10000)
pl = BatchPlanner(bench.get_lattice(tag), device="cuda:0", stateful=True)
pl.set_vel_params(**bench.vel_kwargs())
pl.stage_scenarios(sc); pl.upload(); pl.set_startpos(); pl.tick()
pos_e, vel_e = sc.pos.copy(), sc.vel.copy()
for k in range(3):
    out = pl.download()
    rows, lens, acts = out["traj_row"].numpy(), out["traj_len"].numpy(), out["action_id"].numpy()
    slot = np.argmax(rows >= 0, axis=0); bidx = np.arange(rows.shape[1]); ok = rows[slot, bidx] >= 0
    r = np.where(ok, rows[slot, bidx], 0)
    tr = out["traj"].numpy()[r].astype(np.float64); n = np.maximum(lens[slot, bidx], 2)
    s_t = tr[:, 0, 0] + np.maximum(tr[:, 0, 5] * dt + 0.5 * tr[:, 0, 6] * dt ** 2, 0.0)
    valid = np.arange(tr.shape[1])[None, :] < n[:, None]
    i0 = np.clip((np.where(valid, tr[:, :, 0], np.inf) <= s_t[:, None]).sum(axis=1) - 1, 0, n - 2)
    s0, s1 = tr[bidx, i0, 0], tr[bidx, i0 + 1, 0]
    f = np.clip((s_t - s0) / np.maximum(s1 - s0, 1e-9), 0.0, 1.0)
    lerp = lambda c: tr[bidx, i0, c] * (1 - f) + tr[bidx, i0 + 1, c] * f
    pos_e = np.where(ok[:, None], np.column_stack((lerp(1), lerp(2))), pos_e); vel_e = np.where(ok, lerp(5), vel_e)
    sel = np.where(ok, acts[slot, bidx], 0).astype(np.int32)
    pl.next_tick(ScenarioBatch(pos_e.copy(), sc.heading, sc.vel, sc.n_obj, sc.obj), sel, 2.0 * dt, vel_est=vel_e)
torch.cuda.synchronize()
