#!/usr/bin/env python
"""
bench.py -- planning ticks/s of the batched online planning path (BASELINE.json metric) on N B200 GPUs.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (CPU arm: oracle port on all host cores)

A "step" = one planning tick (calc_paths + calc_vel_profile, 4 kernels) over one batch of 10 000 synthetic scenarios
(Monteblanco lattice with lat_resolution=1.0, lon_straight_step=12.0 -> 216 layers x 7..12 nodes;
random ego arc length + 1..3 dynamic obstacles, seed 20260924).  set_startpos is setup.
  value : device-timed (CUDA events on the launching stream), inputs resident in HBM, L2 flushed between steps
  e2e   : same metric through the public API Graph_LTPL.plan_stream with HOST buffers: per step host staging + H2D of
          the scenario arrays, set_startpos + tick kernels, D2H of the per-path arrays and of the kept trajectory rows,
          all inside the timed region (the D2H of step i overlaps the kernels of step i + 1)
  roofline / cpu_baseline : see DESIGN.md "Measurement"
Multi-GPU: scenarios are independent -> every rank plans its own 10 000-scenario batch ("weak" scaling), no data-path
collective; the lattice blob is NCCL-broadcast from rank 0 at init and the e2e leg all-gathers the action sets.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
# stdout carries exactly ONE JSON line.  NCCL prints its version banner (and any NCCL_DEBUG output) to the process's
# stdout, so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate of the
# original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


import numpy as np  # noqa: E402

TRACK_CSV = os.path.join(REPO, "inputs", "traj_ltpl_cl", "traj_ltpl_cl_monteblanco.csv")
OFFLINE_INI = os.path.join(REPO, "params", "ltpl_config_offline.ini")
ONLINE_INI = os.path.join(REPO, "params", "ltpl_config_online.ini")
AXM_CSV = os.path.join(REPO, "inputs", "veh_dyn_info", "ax_max_machines.csv")
LATTICES = {"l216": {"lat_resolution": 1.0, "lon_straight_step": 12.0}, "default": {},
            "l430": {"lon_curve_step": 6.0, "lon_straight_step": 6.0, "lat_resolution": 0.5}}
WORKLOADS = {"l216": "config2: 10k scenarios x (ego start + 1-3 dynamic obstacles), Monteblanco 216 layers x 7-12 nodes "
                     "(lat_resolution=1.0, lon_straight_step=12.0)",
             "default": "10k scenarios x (ego start + 1-3 dynamic obstacles), Monteblanco shipped ini 128 layers x 13-25",
             "l430": "config4: 10k scenarios x 5 dynamic obstacles, Monteblanco 430 layers x 13-25 nodes"}
METRIC = "planning ticks/s (10k-scenario batch, ~200x11 lattice)"
SEED = 20260924


def ax_max_machines():
    tab = np.loadtxt(AXM_CSV, comments='#', delimiter=',')
    return np.vstack((tab, [100.0, tab[-1, 1]]))


def vel_kwargs():
    return dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=ax_max_machines(), safety_d=30.0)


def lattice_cache_path(tag):
    d = os.path.join(REPO, ".lattice_cache")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "lattice_%s.npz" % tag)


def get_lattice(tag):
    from graphbasedlocaltrajectoryplanner_b200.lattice import load_or_build_lattice
    lat, _ = load_or_build_lattice(TRACK_CSV, OFFLINE_INI, store_path=lattice_cache_path(tag), overrides=LATTICES[tag])
    return lat


def make_batch(tag, batch, seed=SEED):
    from graphbasedlocaltrajectoryplanner_b200.scenarios import Track, make_scenarios
    omin, omax = (5, 5) if tag == "l430" else (1, 3)
    return make_scenarios(Track(TRACK_CSV), batch, seed=seed, n_obj_min=omin, n_obj_max=omax)


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/ltpl_oracle.py) on all host cores; only this leg may execute oracle/
# ----------------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(lat_path):
    os.environ['OPENBLAS_NUM_THREADS'] = '1'   # as main_min_example.py:8
    from graphbasedlocaltrajectoryplanner_b200.lattice import Lattice
    from oracle.ltpl_oracle import OracleLTPL
    _W['orc'] = OracleLTPL(Lattice.load(lat_path))
    _W['vk'] = vel_kwargs()


def _cpu_work(args):
    pos, heading, vel, objs = args
    orc, vk = _W['orc'], _W['vk']
    n_traj = 0
    for i in range(pos.shape[0]):
        ol = [{'id': k + 1, 'type': 'physical', 'X': o[0], 'Y': o[1], 'theta': o[2], 'v': o[3], 'length': o[4],
               'width': 2.5} for k, o in enumerate(objs[i])]
        r = orc.tick(pos[i], heading[i], vel[i], ol, vk)
        n_traj += len(r.get('traj', {}))
    return pos.shape[0], n_traj


class CpuArm(object):
    def __init__(self, tag, cores=None):
        import multiprocessing as mp
        get_lattice(tag)   # make sure the cache file exists
        self.cores = cores or os.cpu_count()
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init, initargs=(lattice_cache_path(tag),))
        self.pool.map(_cpu_work, [self._chunk(make_batch(tag, self.cores, seed=1), i, i + 1)
                                  for i in range(self.cores)])   # start-up + first-touch, untimed

    @staticmethod
    def _chunk(sc, a, b):
        return (sc.pos[a:b], sc.heading[a:b], sc.vel[a:b],
                [sc.obj[i, :sc.n_obj[i]].tolist() for i in range(a, b)])

    def run(self, sc):
        n = sc.size
        per = max(1, n // (self.cores * 4))
        chunks = [self._chunk(sc, a, min(a + per, n)) for a in range(0, n, per)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_work, chunks)
        dt = time.perf_counter() - t0
        assert sum(r[0] for r in res) == n
        return dt

    def close(self):
        self.pool.close()
        self.pool.join()


# ----------------------------------------------------------------------------------------------------------------------
def clocks_sampler(gpu_index, stop_evt, out, ready_evt=None):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20", "-i",
                              str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        if ready_evt is not None:
            ready_evt.set()
        return
    lines = []

    def reader():
        for line in p.stdout:
            lines.append(line)
            if ready_evt is not None:
                ready_evt.set()
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    stop_evt.wait()
    p.terminate()
    th.join(timeout=2)
    sm, mx, reasons = [], [], set()
    for line in lines:
        f = [x.strip() for x in line.split(",")]
        if len(f) < 6:
            continue
        try:
            sm.append(float(f[0]))
            mx.append(float(f[1]))
        except ValueError:
            continue
        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
            if val.lower().startswith("active"):
                reasons.add(name)
    if sm:
        out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def algorithmic_bytes(lat, stats):
    """ALGORITHMIC bytes per launch of each kernel (every logical array element touched once; float64 / int32 as the
    kernels use them).  Formulas: DESIGN.md "Kernels and rooflines"."""
    n_pairs = max(1, lat.num_layers)
    e_l = lat.num_edges / n_pairs                     # edges per layer pair
    s_l = lat.num_samples / n_pairs                   # samples per layer pair
    n_l = lat.num_nodes / lat.num_layers
    discs = 2.0 * stats["n_obj_sum"]                  # object + 0.2 s prediction (OLI:121-127)
    ah = stats["seg_sum"]                             # sum over found actions of path segments (H)
    pts = stats["pts_sum"]                            # sum over found actions of path points (P)
    pts_follow = stats["pts_follow_sum"]
    acts = stats["n_actions"]
    B = stats["batch"]
    plan = discs * 2 * s_l * (16 + 4) + ah * e_l * (8 + 4 + 0.125) + ah * n_l * 1 + acts * 0 \
        + stats["nodes_sum"] * (8 + 4) + B * (16 + 8 + 8 + 4) + stats["n_obj_sum"] * 40 + B * stats["p0_mean"] * 16
    path = pts * (5 * 8 + 8) + ah * (8 * 8 + 16 + 8 + 8 + 4) + acts * stats["p0_mean"] * 5 * 8 * 2
    vel = pts * (2 * 8 + 3 * 8) + pts_follow * (2 * 8 + 3 * 8 * 2)
    export = stats["export_rows"] * (7 * 8 + 7 * 4)
    return dict(k_plan=plan, k_path=path, k_vel=vel, k_export=export)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=10000, help="scenarios per GPU")
    ap.add_argument("--lattice", default="l216", choices=sorted(LATTICES))
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra default-lattice / velocity microbench lines")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:   # convenience: re-exec under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    tag = args.lattice
    config = {"workload": WORKLOADS[tag], "lattice": tag, "per_gpu_batch": args.batch, "seed": SEED,
              "stateless_first_tick": True, "velocity_planner": "fb"}

    # ------------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count()
        sample = args.cpu_sample or max(cores * 8, 512)
        arm = CpuArm(tag, cores)
        sc = make_batch(tag, sample)
        for _ in range(args.warmup):
            arm.run(sc.subset(np.arange(min(sample, cores * 2))))
        t = [arm.run(sc) for _ in range(args.steps)]
        arm.close()
        total = float(np.sum(t))
        value = sample * args.steps / total
        desc = "first %d scenarios of the seeded batch per step" % sample
        emit(({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "ticks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config, cpu_sample=sample),
            "cpu_baseline": {"value": value, "unit": "ticks/s", "cores": cores, "kind": "port", "sample": desc,
                             "note": "float64 NumPy restatement of the reference path (oracle/ltpl_oracle.py); the "
                                     "reference itself needs igraph + tph which are not installable offline"},
            "e2e": {"value": value, "unit": "ticks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:   # before CUDA is initialised in this process (spawned workers)
        cores = os.cpu_count()
        sample = args.cpu_sample or max(cores * 8, 512)
        arm = CpuArm(tag, cores)
        sc_cpu = make_batch(tag, sample)
        arm.run(sc_cpu.subset(np.arange(min(sample, cores * 2))))
        dt = min(arm.run(sc_cpu) for _ in range(2))
        arm.close()
        cpu = {"value": sample / dt, "unit": "ticks/s", "cores": cores, "kind": "port",
               "sample": "first %d scenarios of the seeded batch, best of 2, %d worker processes" % (sample, cores)}

    import torch
    import torch.distributed as dist
    from graphbasedlocaltrajectoryplanner_b200 import capi, parallel
    from graphbasedlocaltrajectoryplanner_b200.Graph_LTPL import Graph_LTPL
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner, read_online_config

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # lattice: built on rank 0, broadcast as one byte blob (NCCL), every rank creates its own handle
    lat = get_lattice(tag) if rank == 0 or world == 1 else None
    if world > 1:
        header, cap, blob_t = parallel.broadcast_lattice(lat, device, src=0)
        pl = BatchPlanner(online=read_online_config(ONLINE_INI), device=device, packed=(header, cap), blob_tensor=blob_t)
        if lat is None:
            lat = get_lattice(tag)   # host-side stats only
    else:
        pl = BatchPlanner(lat, online=read_online_config(ONLINE_INI), device=device)
    pl.set_vel_params(**vel_kwargs())

    # every rank plans its own batch (different seeds per rank): weak scaling, no data-path collective
    sc = make_batch(tag, args.batch, seed=SEED + rank)
    pl.stage_scenarios(sc)
    pl.upload()
    pl.set_startpos()
    torch.cuda.synchronize(device)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2
    stream = torch.cuda.current_stream(device)

    def timed_loop(fn, steps):
        """sum of per-step device times (CUDA events on the launching stream), L2 flushed before every step."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize(device)
        return sum(e0.elapsed_time(e1) for e0, e1 in evs) * 1e-3

    for _ in range(args.warmup):
        pl.tick()
    clocks = {}
    stop_evt = threading.Event()
    ready_evt = threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(local_rank, stop_evt, clocks, ready_evt), daemon=True)
    if rank == 0:
        th.start()
        ready_evt.wait(timeout=10.0)   # first nvidia-smi sample has arrived: the sampler covers the timed region
    barrier()
    l0 = pl.launch_count()
    t_dev = timed_loop(pl.tick, args.steps)
    launches = pl.launch_count() - l0
    barrier()
    t_dev = max_over_ranks(t_dev)
    value = world * args.batch * args.steps / t_dev

    # batch statistics for the algorithmic byte counts (rank 0's batch)
    f = pl.fetch("sc_flags", "status", "action_id", "n_nodes", "path_len", "traj_len", "const_len")
    found = (f["status"] & capi.ST_FOUND) != 0
    stats = dict(batch=args.batch, n_obj_sum=float(sc.n_obj.sum()), n_actions=float(found.sum()),
                 seg_sum=float(np.maximum(f["n_nodes"] - 2, 0)[found].sum()), nodes_sum=float(f["n_nodes"][found].sum()),
                 pts_sum=float(f["path_len"][found].sum()),
                 pts_follow_sum=float(f["path_len"][found & (f["action_id"] == capi.ACT_FOLLOW)].sum()),
                 export_rows=float(f["traj_len"].sum()), p0_mean=float(f["const_len"].mean()))
    n_bad = {name: int(((f["sc_flags"] & bit) != 0).sum()) for name, bit in (
        ("out_of_track", capi.SC_OUT_OF_TRACK), ("heading_mismatch", capi.SC_HEADING_MISMATCH),
        ("capacity", capi.SC_CAPACITY), ("brake_prefix", capi.SC_BRAKE_PREFIX))}

    # per-kernel device time (each kernel launched alone through ltpl_launch_stage, same buffers, L2 flushed)
    import ctypes as C
    stage_names = {1: "k_plan", 2: "k_path", 3: "k_vel", 4: "k_export"}
    ktime = {}
    for stage, name in stage_names.items():
        def one(stage=stage):
            capi.check(pl.lib, pl.lib.ltpl_launch_stage(stage, pl.handle, C.byref(pl.params), C.byref(pl.dims),
                                                        C.byref(pl.buf), pl.stream), name)
        for _ in range(3):
            one()
        ktime[name] = timed_loop(one, args.steps) / args.steps

    # ------------------------------------------------------------------------------------------------------------------
    # e2e: public API with host buffers (pinned): H2D scenario arrays + set_startpos + tick + D2H action sets per step
    ltpl = Graph_LTPL.__new__(Graph_LTPL)
    ltpl._Graph_LTPL__planner = pl           # reuse the planner (same lattice handle / buffers)
    cap_rows = (3 * args.batch) // 2   # fixed-stride all-gather capacity (mean is ~1.3 kept trajectories per scenario)
    comm = torch.cuda.Stream(device=device) if world > 1 else None
    gather_bufs, gather_done = [], {}
    snap = [(torch.empty_like(pl.t["traj_len"]), torch.empty_like(pl.t["traj_id"])) for _ in range(pl.N_SETS)]

    def hook(k):
        # all-gather of the (fixed-stride) compact action sets over NVLink on a communication stream: it overlaps the
        # kernels of the next step exactly like the D2H copy does (the small per-path arrays are snapshotted first)
        if world == 1:
            return
        compute = torch.cuda.current_stream(device)
        snap[k][0].copy_(pl.t["traj_len"])
        snap[k][1].copy_(pl.t["traj_id"])
        ev = torch.cuda.Event()
        ev.record(compute)
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            parallel.gather_action_sets(pl.traj_bufs[k][:cap_rows], snap[k][0], snap[k][1], out=gather_bufs)
            done = torch.cuda.Event()
            done.record(comm)
        gather_done[k] = done

    def hook_before(k):
        if k in gather_done:   # buffer set k is about to be rewritten
            torch.cuda.current_stream(device).wait_event(gather_done[k])
    hook.before = hook_before

    def feed(n):
        for _ in range(n):
            yield sc      # the same host-side ScenarioBatch is staged, uploaded and planned every step

    for out in ltpl.plan_stream(feed(3), device_hook=hook):
        pass
    barrier()
    t0 = time.perf_counter()
    rows = 0
    for out in ltpl.plan_stream(feed(args.steps), device_hook=hook):
        rows += int(out["n_rows"])
    torch.cuda.synchronize(device)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * args.batch * args.steps / t_e2e
    assert rows > 0 and rows <= cap_rows * args.steps
    rows_per_step = rows / args.steps
    if rank == 0:
        stop_evt.set()
        th.join(timeout=3)

    # third timing of SURVEY 8(d): the optional per-scenario Python view of one result (reference-style dicts)
    t0 = time.perf_counter()
    unpacked = Graph_LTPL.unpack_batch(out)
    t_unpack = time.perf_counter() - t0
    assert len(unpacked) == args.batch

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    alg = algorithmic_bytes(lat, stats)
    dom = max(ktime, key=ktime.get)
    achieved = alg[dom] / ktime[dom] / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom],
                "kernel_ms": {k: 1e3 * v for k, v in ktime.items()},
                "kernel_share_of_step": {k: v / sum(ktime.values()) for k, v in ktime.items()},
                "per_kernel_gbs": {k: alg[k] / ktime[k] / 1e9 for k in ktime},
                "note": "latency/issue-bound: the lattice (%.1f MB) is L2-resident, see DESIGN.md" % (
                    pl.blob.numel() / 1e6)}
    traffic_path = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.isfile(traffic_path):
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get(tag, {}).get(dom)
        except Exception:
            pass

    result = {
        "metric": METRIC, "value": value, "unit": "ticks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config, l2="256 MiB buffer written before every timed step (L2 = 126 MB)",
                       actions_per_tick=stats["n_actions"] / args.batch,
                       path_points_per_action=stats["pts_sum"] / max(stats["n_actions"], 1),
                       scenarios_flagged=n_bad, parallelism="%d x independent scenario shards" % world),
        "e2e": {"value": e2e_value, "unit": "ticks/s", "h2d_bytes_per_step": pl.h2d_bytes(),
                "d2h_bytes_per_step": pl.d2h_bytes(int(rows_per_step)), "ms_per_step": 1e3 * t_e2e / args.steps,
                "kept_trajectories_per_step": rows_per_step,
                "facade_unpack_ms_per_batch": 1e3 * t_unpack,
                "api": "Graph_LTPL.plan_stream: per step host staging + H2D + set_startpos + calc_paths + "
                       "calc_vel_profile + D2H of the compact action sets; D2H of step i overlaps the kernels of step "
                       "i+1 (copy stream, 3 buffer sets)" + ("; + all_gather of the action sets on a communication stream" if world > 1 else "")},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cpu is not None:
        result["cpu_baseline"] = cpu

    if not args.no_extra and world == 1:
        extra = {}
        try:   # shipped default lattice (128 layers x 13-25 nodes, 14 k edges): the non-degenerate search workload
            lat_d = get_lattice("default")
            pl_d = BatchPlanner(lat_d, online=read_online_config(ONLINE_INI), device=device)
            pl_d.set_vel_params(**vel_kwargs())
            pl_d.stage_scenarios(make_batch("default", args.batch))
            pl_d.upload()
            pl_d.set_startpos()
            for _ in range(args.warmup):
                pl_d.tick()
            td = timed_loop(pl_d.tick, args.steps)
            extra["default_lattice_ticks_per_s"] = args.batch * args.steps / td
            extra["default_lattice_ms_per_step"] = 1e3 * td / args.steps
            del pl_d
        except Exception as e:   # noqa: BLE001
            extra["default_lattice_error"] = str(e)[:200]
        try:   # 430 layers x 13-25 nodes (lon steps 6 m, lat_resolution 0.5), 5 objects each
            pl_4 = BatchPlanner(get_lattice("l430"), online=read_online_config(ONLINE_INI), device=device)
            pl_4.set_vel_params(**vel_kwargs())
            pl_4.stage_scenarios(make_batch("l430", args.batch))
            pl_4.upload()
            pl_4.set_startpos()
            for _ in range(args.warmup):
                pl_4.tick()
            t4 = timed_loop(pl_4.tick, args.steps)
            extra["config4_l430_ticks_per_s"] = args.batch * args.steps / t4
            extra["config4_l430_ms_per_step"] = 1e3 * t4 / args.steps
            del pl_4
        except Exception as e:   # noqa: BLE001
            extra["config4_l430_error"] = str(e)[:200]
        try:   # stateful ticks: closed loop of 8 ticks on the bench workload; a
            # vehicle dummy advances every scenario 0.1 s on its first kept trajectory; the loop is recorded once
            # (untimed host work between the ticks) and replayed with CUDA events around every next_tick
            from graphbasedlocaltrajectoryplanner_b200.scenarios import ScenarioBatch
            pl_s = BatchPlanner(lat, online=read_online_config(ONLINE_INI), device=device, stateful=True)
            pl_s.set_vel_params(**vel_kwargs())
            n_loop, dt_loop = 8, 0.1

            def first_tick():
                pl_s.stage_scenarios(sc)
                pl_s.upload()
                pl_s.set_startpos()
                pl_s.tick()

            def advance(out):
                rows = out["traj_row"].numpy()
                lens = out["traj_len"].numpy()
                acts = out["action_id"].numpy()
                slot = np.argmax(rows >= 0, axis=0)                      # first kept trajectory of every scenario
                bidx = np.arange(rows.shape[1])
                ok = rows[slot, bidx] >= 0
                r = np.where(ok, rows[slot, bidx], 0)
                tr = out["traj"].numpy()[r].astype(np.float64)           # (B, 115, 7)
                n = np.maximum(lens[slot, bidx], 2)
                s_t = tr[:, 0, 0] + np.maximum(tr[:, 0, 5] * dt_loop + 0.5 * tr[:, 0, 6] * dt_loop ** 2, 0.0)
                valid = np.arange(tr.shape[1])[None, :] < n[:, None]
                i0 = np.clip((np.where(valid, tr[:, :, 0], np.inf) <= s_t[:, None]).sum(axis=1) - 1, 0, n - 2)
                s0, s1 = tr[bidx, i0, 0], tr[bidx, i0 + 1, 0]
                f = np.clip((s_t - s0) / np.maximum(s1 - s0, 1e-9), 0.0, 1.0)
                lerp = lambda c: tr[bidx, i0, c] * (1 - f) + tr[bidx, i0 + 1, c] * f   # noqa: E731
                return np.column_stack((lerp(1), lerp(2))), lerp(5), np.where(ok, acts[slot, bidx], 0), ok

            first_tick()
            rec_in = []
            pos_e, vel_e = sc.pos.copy(), sc.vel.copy()
            for _ in range(n_loop):
                out = pl_s.download()
                torch.cuda.synchronize(device)
                p_new, v_new, sel_a, ok = advance(out)
                pos_e, vel_e = np.where(ok[:, None], p_new, pos_e), np.where(ok, v_new, vel_e)
                rec_in.append((pos_e.copy(), vel_e.copy(), sel_a.astype(np.int32)))
                sc_k = ScenarioBatch(pos_e.copy(), sc.heading, sc.vel, sc.n_obj, sc.obj)
                pl_s.next_tick(sc_k, sel_a, 2.0 * dt_loop, vel_est=vel_e)
            torch.cuda.synchronize(device)
            flags = pl_s.fetch("sc_flags")["sc_flags"]
            first_tick()
            t_st = 0.0
            stream_s = torch.cuda.current_stream(device)
            for pos_k, vel_k, sel_k in rec_in:
                sc_k = ScenarioBatch(pos_k, sc.heading, sc.vel, sc.n_obj, sc.obj)
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream_s)
                pl_s.next_tick(sc_k, sel_k, 2.0 * dt_loop, vel_est=vel_k)
                e1.record(stream_s)
                torch.cuda.synchronize(device)
                t_st += e0.elapsed_time(e1) * 1e-3
            extra["stateful_tick"] = {"ticks_per_s": args.batch * n_loop / t_st, "ms_per_tick": 1e3 * t_st / n_loop,
                                      "ticks": n_loop, "scenarios_still_planned_at_the_end": int((flags == 0).sum()),
                                      "flags_at_the_end": {name: int(((flags & bit) != 0).sum()) for name, bit in (
                                          ("out_of_track", capi.SC_OUT_OF_TRACK),
                                          ("heading_mismatch", capi.SC_HEADING_MISMATCH), ("capacity", capi.SC_CAPACITY),
                                          ("brake_prefix", capi.SC_BRAKE_PREFIX),
                                          ("state_fallback", capi.SC_STATE_FALLBACK))},
                                      "note": "closed loop, 0.1 s per tick; device time of one next_tick incl. its input upload"}
            del pl_s
        except Exception as e:   # noqa: BLE001
            extra["stateful_tick_error"] = str(e)[:300]
        try:   # 100 k paths x 500 points forward/backward solver
            from graphbasedlocaltrajectoryplanner_b200.scenarios import make_velocity_microbench
            from graphbasedlocaltrajectoryplanner_b200.velprofile import velprofile_batch_device
            mb = make_velocity_microbench(100000, 500)
            d = {k: torch.from_numpy(np.ascontiguousarray(mb[k])).to(device) for k in ("kappa", "el", "v_start", "v_end")}
            vx = torch.empty_like(d["kappa"])
            ax = torch.empty_like(d["kappa"])
            pl.set_vel_params(vel_max=60.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=ax_max_machines(),
                              safety_d=30.0)

            def vp():
                velprofile_batch_device(pl, d["kappa"], d["el"], d["v_start"], d["v_end"], vx, ax)
            for _ in range(3):
                vp()
            tv = timed_loop(vp, 10) / 10
            nbytes = 100000 * 500 * 4 * 8
            extra["velprofile_100k_x_500"] = {"ms": 1e3 * tv, "paths_per_s": 100000 / tv,
                                              "algorithmic_GBps": nbytes / tv / 1e9, "frac_of_peak": nbytes / tv / 1e9 / peak,
                                              "dtype": "f64 (kappa, el in; vx, ax out = 32 B / point)"}
        except Exception as e:   # noqa: BLE001
            extra["velprofile_error"] = str(e)[:200]
        result["extra"] = extra

    emit(result)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
