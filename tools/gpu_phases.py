"""debug (GPU box): build an instrumented library and print average cycles per phase of k_vel."""
import ctypes, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from graphbasedlocaltrajectoryplanner_b200 import capi
dbg = "/tmp/libltpl_dbg.so"
subprocess.check_call(["nvcc"] + capi.NVCC_FLAGS + ["-DLTPL_PROFILE_PHASES", "-I" + capi.INCLUDE_DIR, "-o", dbg,
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
nw = (cnt[0] + 7) // 8 + (cnt[1] + 7) // 8
names = {0: "pass A (s, ego brake, nearest)", 1: "follow scalars", 2: "follow sweeps (fwd+bwd)", 3: "reduced merge", 4: "single profile (fwd+bwd)", 5: "output pass"} if os.environ.get("TILED", "1") == "1" else {0: "cumsum s", 1: "2x s_coord on path", 2: "ego brake", 3: "glob_rl match", 4: "opp brake+stop idx+vctrl",
         5: "control profile", 6: "complete profile", 7: "min", 8: "non-follow fb / red", 9: "ax+sqrt", 10: "follow total tail"}
print("queue counts", cnt[:2], "warps", nw)
pn = {16: "plan: defaults+object filter", 17: "plan: planning range", 18: "plan: blocked edges+closest", 19: "plan: const-seg objects",
      20: "plan: glob match", 21: "plan: DP", 22: "plan: goal+backtrack"}
for k in range(16, 23):
    print("%-30s %10.0f cycles/scenario" % (pn[k], out[k] / 10000.0))
tot = sum(out[:16])
for k in range(6):
    print("%-28s %12.0f cycles/warp-with-phase(avg over all warps) %5.1f%%" % (names.get(k, k), out[k] / nw, 100.0 * out[k] / max(tot, 1)))
