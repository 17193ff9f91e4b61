"""large parity sweep (GPU box): CUDA path vs the oracle over thousands of seeded scenarios per lattice, compared with the
rules of tests/helpers.compare_records (node sequences bit-exact, coordinates / velocities 1e-4).  The oracle runs in a
process pool on the host cores.   python tools/gpu_parity_sweep.py [n_per_lattice] [lattice ...]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np

_W = {}


def _init(tag):
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    from tests import helpers as H
    from oracle.ltpl_oracle import OracleLTPL
    import bench
    _W['H'] = H
    _W['orc'] = OracleLTPL(bench.get_lattice(tag))
    _W['vk'] = bench.vel_kwargs()


def _work(args):
    idx, pos, heading, vel, ols, recs = args
    H, orc, vk = _W['H'], _W['orc'], _W['vk']
    bad, ties, acts = [], 0, 0
    for i in range(len(idx)):
        want = orc.tick(pos[i], heading[i], vel[i], ols[i], vk)
        got = recs[i]
        try:
            H.compare_records(got, want, ctx="scenario %d" % idx[i])
        except AssertionError as e:
            bad.append(str(e).split("\n")[0][:300])
        if not want["out_of_track"]:
            acts += len(want["paths"])
            ties += sum(1 for a in want.get("tie", {}) if want["tie"][a] or got.get("tie", {}).get(a))
    return bad, ties, acts


def main():
    import multiprocessing as mp
    import bench
    from graphbasedlocaltrajectoryplanner_b200.planner import BatchPlanner
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    tags = sys.argv[2:] or ["l216", "default", "l430"]
    total_bad = 0
    for tag in tags:
        sc = bench.make_batch(tag, n, seed=bench.SEED + 17)
        pl = BatchPlanner(bench.get_lattice(tag), device="cuda:0")
        pl.set_vel_params(**bench.vel_kwargs())
        pl.stage_scenarios(sc); pl.upload(); pl.set_startpos(); pl.tick()
        t0 = time.time()
        recs = pl.records()
        chunks = []
        step = 40
        for a in range(0, n, step):
            b = min(a + step, n)
            chunks.append((list(range(a, b)), sc.pos[a:b], sc.heading[a:b], sc.vel[a:b],
                           [sc.object_list(i) for i in range(a, b)], recs[a:b]))
        ctx = mp.get_context("spawn")
        with ctx.Pool(min(os.cpu_count(), 96), initializer=_init, initargs=(tag,)) as pool:
            res = pool.map(_work, chunks)
        bad = [m for r in res for m in r[0]]
        ties = sum(r[1] for r in res)
        acts = sum(r[2] for r in res)
        total_bad += len(bad)
        print("[%s] %d scenarios, %d action paths compared, %d tie-flagged (skipped), %d mismatches  (%.1f s)" %
              (tag, n, acts, ties, len(bad), time.time() - t0))
        for m in bad[:8]:
            print("   ", m)
    print("TOTAL mismatches:", total_bad)


if __name__ == "__main__":
    main()
