"""Single-core rate of the UNMODIFIED reference (graph_ltpl from /root/reference) on the bench workload, next to the
repo's NumPy port, measured in the build container (the reference checkout does not exist on the GPU box, and its two
third-party dependencies are restated by oracle/shims -- kind "reference+shims").  Writes one JSON object:

    python oracle/time_reference_shims.py [n_scenarios] > profiles/r2_cpu_reference_shims.json

Per scenario, as BASELINE.md section 2 asks: set_startpos -> calc_paths -> calc_vel_profile, visual_mode=False,
log_to_file=False, one BLAS thread (main_min_example.py:8)."""
import json
import os
import sys
import time

for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[k] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
real_stdout = os.dup(1)   # bench.py points fd 1 at stderr when it is imported
import numpy as np  # noqa: E402

import bench  # noqa: E402
from oracle import gen_golden as G  # noqa: E402
from oracle.ltpl_oracle import OracleLTPL  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
tag = "l216"
os.dup2(2, 1)   # the reference prints progress bars
graph_ltpl = G.load_reference()
ltpl, _ = G.make_ltpl(graph_ltpl, tag, bench.LATTICES[tag])
vk = dict(bench.vel_kwargs(), incl_emerg_traj=False)
sc = bench.make_batch(tag, n)
for b in range(4):   # warm-up
    G.run_tick(ltpl, sc.pos[b], sc.heading[b], sc.vel[b], sc.object_list(b), vk)
t0 = time.perf_counter()
ok = 0
for b in range(n):
    r = G.run_tick(ltpl, sc.pos[b], sc.heading[b], sc.vel[b], sc.object_list(b), vk)
    ok += int(not r["out_of_track"])
t_ref = time.perf_counter() - t0
orc = OracleLTPL(bench.get_lattice(tag))
vk_o = {k: v for k, v in vk.items() if k != "incl_emerg_traj"}
for b in range(4):
    orc.tick(sc.pos[b], sc.heading[b], sc.vel[b], sc.object_list(b), vk_o)
t0 = time.perf_counter()
for b in range(n):
    orc.tick(sc.pos[b], sc.heading[b], sc.vel[b], sc.object_list(b), vk_o)
t_port = time.perf_counter() - t0
sys.stdout.flush()
os.dup2(real_stdout, 1)
print(json.dumps({
    "what": "single-core planning ticks/s on the bench workload (first %d scenarios of the seeded l216 batch)" % n,
    "kind_reference": "reference+shims: the unmodified graph_ltpl of /root/reference @ 18763ef9 on oracle/shims "
                      "(igraph = minimal pure-Python graph class, trajectory_planning_helpers = oracle/tph_port.py)",
    "ticks_per_s_reference_shims_1core": n / t_ref, "ms_per_tick_reference_shims": 1e3 * t_ref / n,
    "kind_port": "oracle/ltpl_oracle.py (the CPU arm of bench.py)",
    "ticks_per_s_port_1core": n / t_port, "ms_per_tick_port": 1e3 * t_port / n,
    "scenarios": n, "planned": ok, "host": "build container (%d usable cores), not the GPU box" % bench.usable_cores(),
    "note": "the igraph shim is pure Python where python-igraph is C: graph copies / searches of the reference run "
            "slower here than on a real install; the port is the faster and therefore the stricter CPU baseline"}))
