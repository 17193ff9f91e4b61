"""debug (GPU box): build an instrumented library and print average cycles per phase of k_vel."""
import ctypes, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from graphbasedlocaltrajectoryplanner_b200 import capi
dbg = "/tmp/libltpl_dbg.so"
subprocess.check_call(["nvcc"] + capi.NVCC_FLAGS + ["-DLTPL_PROFILE_PHASES", "-DVR_P=" + os.environ.get("VR_P", "8"), "-I" + capi.INCLUDE_DIR, "-o", dbg,
                       os.path.join(capi.CSRC_DIR, "ltpl_api.cu")])
capi.LIB_PATH = dbg
import numpy as np, torch
import bench
from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
tag = sys.argv[1] if len(sys.argv) > 1 else "l216"
pl = BatchPlanner(bench.get_lattice(tag), device="cuda:0")
pl.set_vel_params(**bench.vel_kwargs())
pl.stage_scenarios(bench.make_batch(tag, 10000)); pl.upload(); pl.set_startpos()
for _ in range(3): pl.tick()
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
pl.lib.ltpl_debug_phases(None, 1)
pl.tick(); torch.cuda.synchronize()
pl.lib.ltpl_debug_phases(out, 0)
cnt = pl.t["queue_cnt"].cpu().numpy()
P = int(os.environ.get("VR_P", "8"))
gf = (cnt[0] + P - 1) // P
# time since the previous marker of the same warp, summed over both warps of every CTA (all classes)
names = {6: "setup + (w1) opponent brake distance", 0: "wait for the bulk copies", 1: "rows -> fp32 (w0: el, s scan; w1: kappa)",
         8: "w1 nearest path points", 7: "w1 up to the brake chain", 9: "w1 wait for E2 / arc lengths (bar 2)",
         10: "w1 ego brake chain + follow scalars", 11: "w1 control profile fwd + bwd, acceptance",
         3: "w0 forward sweep (complete profile)", 4: "w0 backward sweep", 5: "w0 wait for warp 1", 12: "w1 wait for warp 0",
         13: "element-wise end + export rows"}
print("queue counts", cnt[:2], "follow CTAs", gf)
pn = {16: "plan: defaults+object filter", 17: "plan: planning range", 18: "plan: blocked edges+closest", 19: "plan: const-seg objects",
      20: "plan: glob match", 21: "plan: DP", 22: "plan: goal+backtrack"}
for k in range(16, 23):
    print("%-30s %10.0f cycles/scenario" % (pn[k], out[k] / 10000.0))
for k in (6, 0, 1, 8, 7, 9, 10, 11, 3, 4, 5, 12, 13):
    print("%-44s %10.0f cycles per follow CTA" % (names[k], out[k] / max(gf, 1)))
