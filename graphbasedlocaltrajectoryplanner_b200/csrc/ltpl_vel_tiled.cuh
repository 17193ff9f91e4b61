// ltpl_vel_tiled.cuh -- k_vel_tiled: OTH.calc_vel_profile per action (OTH:688-1025) as TILE-STREAMED SIMT recurrences.
//
// The velocity solver is a serial recurrence per path and a 10 k-scenario batch only holds ~13 k paths, so the kernel
// is bound by (instructions on the critical path of one warp) x (dependent-issue latency), not by bandwidth.  Mapping:
//   * one warp = VT_P queued paths of one class (follow / other); every path owns TWO lanes ("roles"): lane pl runs
//     the complete profile / cumulative arc length / ego brake profile, lane pl + VT_P runs the follow-mode control
//     profile / the nearest-point searches / the ax division.  Both roles execute the same instruction stream, so the
//     two independent recurrences of follow mode cost one; small VT_P spreads the few paths over many warps so that
//     every SM sub-partition has several warps to overlap latencies with.
//   * every array moves in TILES of 32 points x VT_P paths: row-major inputs (kappa, el, x, y per path) are copied
//     global -> shared with cp.async, transposing on the fly (one coalesced 256-byte request per path and tile, all
//     requests of a tile in flight together); lanes then read tile[point][path] for 32 recurrence steps;
//     intermediate profiles live in a TRANSPOSED global scratch [array][point][column] that only this warp touches;
//     results return to the row-major planes through a tile.
// Passes (follow): A forward: s = cumsum(el), ego brake profile | nearest points;   B forward: complete | control;
//                  C backward: complete | control, intersection;  D backward: vx = sqrt(w) | ax; row-major output.
// Other class: A = cumsum, B / C = one profile on role 0.  Float64, w = v^2 domain (see ltpl_vel.cuh).
#pragma once
#include "ltpl_vel.cuh"

#ifndef VT_P
#define VT_P 8                       // paths per warp (power of two, 2 * VT_P <= 32)
#endif
#define VT_W (VT_P + 1)              // padded tile row (doubles)
#ifndef VT_UNROLL
#define VT_UNROLL 8                  // unroll factor of the tile move loops (VT_P iterations each, ~20 call sites)
#endif
#define VT_PRAGMA_(x) _Pragma(#x)
#define VT_PRAGMA_UNROLL(n) VT_PRAGMA_(unroll n)
#define VT_TILE (32 * VT_W)          // doubles per tile
#define VT_NTILES 6                  // tiles per warp
#define VT_SMEM_BYTES (VT_NTILES * VT_TILE * 8 + 2 * VT_P * 4 + 2 * VT_P * 8)

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
}

struct WarpCtx {
    int lane;
    int p_max;
    int ntc;          // columns of the transposed scratch
    int col0;         // first column of this warp
    const int* qs;    // smem[VT_P] path id (-1 = idle slot)
    const int* ns;    // smem[VT_P] points per path (0 = idle slot)
    // smem[VT_P] element offset of point 0 of a path inside a row-major plane: q * p_max (+ the cut index of a stateful
    // tick, ltpl_state.cuh) for the INPUT planes, q * p_max (+ the vel_course rows) for the OUTPUT planes
    const long long* in_base;
    const long long* out_base;
};

// tile[k][r] = row_r[p0 + k] for the warp's VT_P paths (row-major per-path array `plane`); lanes = 32 points
__device__ __forceinline__ void tile_load_rows(const WarpCtx& w, double* tile, const double* plane, int p0) {
VT_PRAGMA_UNROLL(VT_UNROLL)
    for (int r = 0; r < VT_P; ++r) {
        if (p0 + w.lane < w.ns[r])
            cp_async8(&tile[w.lane * VT_W + r], plane + w.in_base[r] + p0 + w.lane);
    }
}
// row_r[p0 + k] = tile[k][r]
__device__ __forceinline__ void tile_store_rows(const WarpCtx& w, const double* tile, double* plane, int p0) {
VT_PRAGMA_UNROLL(VT_UNROLL)
    for (int r = 0; r < VT_P; ++r) {
        if (p0 + w.lane < w.ns[r]) plane[w.out_base[r] + p0 + w.lane] = tile[w.lane * VT_W + r];
    }
}
// transposed scratch <-> tile: 32 rows x VT_P columns, element e = it * 32 + lane -> (row e / VT_P, column e % VT_P)
__device__ __forceinline__ void tile_load_t(const WarpCtx& w, double* tile, const double* tarr, int p0, int np) {
VT_PRAGMA_UNROLL(VT_UNROLL)
    for (int it = 0; it < VT_P; ++it) {
        const int e = it * 32 + w.lane, k = e / VT_P, cc = e % VT_P;
        if (p0 + k < np) cp_async8(&tile[k * VT_W + cc], tarr + (size_t)(p0 + k) * w.ntc + w.col0 + cc);
    }
}
__device__ __forceinline__ void tile_store_t(const WarpCtx& w, const double* tile, double* tarr, int p0, int np) {
VT_PRAGMA_UNROLL(VT_UNROLL)
    for (int it = 0; it < VT_P; ++it) {
        const int e = it * 32 + w.lane, k = e / VT_P, cc = e % VT_P;
        if (p0 + k < np) tarr[(size_t)(p0 + k) * w.ntc + w.col0 + cc] = tile[k * VT_W + cc];
    }
}

// ---- recurrence state machines (same arithmetic as fb_profile_w in ltpl_vel.cuh) ----
struct FwdSt {
    double o_prev, cur, k_prev, e_prev;
    double xlo, xhi, x0, f0, sl;   // cached segment of the machine table: axm(v) = f0 + sl * (v - x0) on [xlo, xhi)
    bool prev_rise, active;
    int hint;
};
__device__ __forceinline__ double fwd_init(FwdSt& s, double oraw, double kabs, double e, double wcap, double wmax) {
    const double o = fmin(fmin(oraw, wmax), wcap);
    s.o_prev = o;
    s.cur = o;
    s.k_prev = kabs;
    s.e_prev = e;
    s.prev_rise = false;
    s.active = false;
    s.hint = 0;
    s.xlo = 1.0;   // empty cache interval
    s.xhi = 0.0;
    s.x0 = 0.0;
    s.f0 = 0.0;
    s.sl = 0.0;
    return o;
}
// np.interp on the machine table with the current segment cached in registers: v moves slowly along a path, so the
// shared-memory search only runs when v leaves [xlo, xhi)
__device__ __forceinline__ double axm_cached(FwdSt& s, double v, const VelCfg& c) {
    if (!(v >= s.xlo && v < s.xhi)) {
        const int n = c.n_axm;
        if (v <= c.axm_v[0]) {
            s.xlo = -LTPL_INF;
            s.xhi = c.axm_v[0];
            s.x0 = 0.0;
            s.f0 = c.axm_a[0];
            s.sl = 0.0;
            if (v == s.xhi) return s.f0;
        } else if (v >= c.axm_v[n - 1]) {
            s.xlo = c.axm_v[n - 1];
            s.xhi = LTPL_INF;
            s.x0 = 0.0;
            s.f0 = c.axm_a[n - 1];
            s.sl = 0.0;
        } else {
            int j = s.hint;
            while (j < n - 2 && v >= c.axm_v[j + 1]) ++j;
            while (j > 0 && v < c.axm_v[j]) --j;
            s.hint = j;
            s.xlo = c.axm_v[j];
            s.xhi = c.axm_v[j + 1];
            s.x0 = s.xlo;
            s.f0 = c.axm_a[j];
            s.sl = c.axm_s[j];
        }
    }
    return fma(s.sl, v - s.x0, s.f0);
}
__device__ __forceinline__ double fwd_step(FwdSt& s, double oraw, double kabs, double e, double wmax, const VelCfg& c) {
    const double o_n = fmin(oraw, wmax);
    const bool rise = o_n > s.o_prev;
    if (!s.active && rise && !s.prev_rise) s.active = true;
    double nxt = o_n;
    if (s.active) {
        // mode 'accel_forw': min(tyre, machine(v)) + drag (acc_forw of ltpl_vel.cuh with the cached table segment)
        double a = acc_tire(s.cur, s.k_prev, c.ax_max, c.inv_ay, c.exp_);
        a = fmin(a, axm_cached(s, sqrt(s.cur), c));
        a = fma(-s.cur, c.dm, a);
        const double wn = fma(2.0 * a, s.e_prev, s.cur);
        if (wn < o_n) nxt = wn;
        if (wn > wmax) s.active = false;
    }
    s.cur = nxt;
    s.prev_rise = rise;
    s.o_prev = o_n;
    s.k_prev = kabs;
    s.e_prev = e;
    return nxt;
}
struct BwdSt {
    double o_prev, cur, k_p;
    bool prev_rise, active;
};
__device__ __forceinline__ void bwd_init(BwdSt& s, double w_hi, double kabs_hi) {
    s.o_prev = w_hi;
    s.cur = w_hi;
    s.k_p = kabs_hi;
    s.prev_rise = false;
    s.active = false;
}
__device__ __forceinline__ double bwd_step(BwdSt& s, double w_p, double kabs_p, double e_p, double wmax,
                                           const VelCfg& c) {
    const double o_n = w_p;
    const bool rise = o_n > s.o_prev;
    if (!s.active && rise && !s.prev_rise) s.active = true;
    double nxt = o_n;
    if (s.active) {
        const double a = acc_backw(s.cur, s.k_p, c);
        double wn = fma(2.0 * a, e_p, s.cur);
        const double a2 = acc_backw(wn, kabs_p, c);
        const double wt = fma(2.0 * a2, e_p, s.cur);
        wn = fmin(wn, wt);
        if (wn < o_n) nxt = wn;
        if (wn > wmax) s.active = false;
    }
    s.cur = nxt;
    s.prev_rise = rise;
    s.o_prev = o_n;
    s.k_p = kabs_p;
    return nxt;
}

// first index i in [0, n) with col[i] >= thr (col non-decreasing), n if none; col = own column of a transposed array
__device__ __forceinline__ int first_ge_t(const double* tcol, int ntc, int n, double thr) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tcol[(size_t)mid * ntc] >= thr)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
// first index i in [0, n) with col[i] <= thr (col non-increasing), n if none
__device__ __forceinline__ int first_le_t(const double* tcol, int ntc, int n, double thr) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tcol[(size_t)mid * ntc] <= thr)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

// get_s_coord.py:8-99 on an open polyline once the nearest index nb is known (x, y, el row-major; s transposed column)
__device__ __forceinline__ double s_coord_from_nb(const double* x, const double* y, const double* el, const double* scol,
                                                  int ntc, int n, int nb, double px, double py) {
    const int idx1 = max(nb - 1, 0), idx2 = min(nb + 1, n - 1);
    const double xn = x[nb], yn = y[nb];
    int ia, ib;
    if (angle_cmp(make_double2(xn, yn), px, py, make_double2(x[idx1], y[idx1]), make_double2(x[idx2], y[idx2])).gt) {
        ia = idx1;
        ib = nb;
    } else {
        ia = nb;
        ib = idx2;
    }
    const bool ins = el[0] > 0.05;   // leading 0 inserted into s_array = cumsum(el) (get_s_coord.py:67-68)
    const double s_ia = scol[(size_t)ia * ntc];
    const double sbase = ins ? s_ia : __dadd_rn(s_ia, el[ia]);
    const double ax = x[ia], ay = y[ia], bx = x[ib] - ax, by = y[ib] - ay;
    const double t = __ddiv_rn(__dadd_rn(__dmul_rn(px - ax, bx), __dmul_rn(py - ay, by)), __dadd_rn(sq_rn(bx), sq_rn(by)));
    const double sx = __dadd_rn(ax, __dmul_rn(t, bx)), sy = __dadd_rn(ay, __dmul_rn(t, by));
    const double ds = sqrt(__dadd_rn(sq_rn(ax - sx), sq_rn(ay - sy)));
    return __dadd_rn(sbase, ds);
}

// One forward + one backward sweep over the tiles for per-lane profiles.  Every COMPUTE lane (lane < 2 VT_P) may own a
// profile on its path's points [lo, hi] (hi < lo: none): start cap wcap, end clamp we (< 0: none), speed limit wmax,
// results into its own tile / transposed array (tsel = 0 / 1 -> tarr0 / tarr1).  zero_tail: zeros behind hi up to n.
// Returns w[lo] after the backward sweep.
__device__ __forceinline__ double profile_sweeps(const WarpCtx& w, double* tiles, const double* kap_pl,
                                                 const double* el_pl, int np, int pl, bool compute, int tsel, int lo,
                                                 int hi, int n, double wcap, double we, double wmax, bool zero_tail,
                                                 const VelCfg& c, double* tarr0, double* tarr1, bool use1) {
    double* t_k = tiles;
    double* t_e = tiles + VT_TILE;
    double* t_w0 = tiles + 2 * VT_TILE;
    double* t_w1 = tiles + 3 * VT_TILE;
    double* t_w = tsel ? t_w1 : t_w0;
    double* t_o = tiles + 4 * VT_TILE;
    const int ntile = (np + 31) >> 5;
    FwdSt f;
    f.cur = 0.0;
    f.hint = 0;
    for (int tl = 0; tl < ntile; ++tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t_k, kap_pl, p0);
        tile_load_rows(w, t_e, el_pl, p0);
        cp_async_wait_all();
        __syncwarp();
        // curvature speed limit w <= ay_max / |kappa| of the whole tile: element-wise, so all 32 lanes share the
        // divisions instead of the 2 VT_P recurrence lanes paying one per step
#pragma unroll
        for (int it = 0; it < VT_P; ++it) {
            const int e = it * 32 + w.lane, k = e / VT_P, cc = e % VT_P;
            t_o[k * VT_W + cc] = c.ay_max / fabs(t_k[k * VT_W + cc]);
        }
        __syncwarp();
        if (compute) {
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const int p = p0 + k;
                if (p >= lo && p <= hi) {
                    const double kabs = fabs(t_k[k * VT_W + pl]);
                    const double e = t_e[k * VT_W + pl];
                    const double oraw = t_o[k * VT_W + pl];
                    double v = (p == lo) ? fwd_init(f, oraw, kabs, e, wcap, wmax) : fwd_step(f, oraw, kabs, e, wmax, c);
                    if (p == hi && we >= 0.0 && v > we) v = we;
                    t_w[k * VT_W + pl] = v;
                } else if (p < n) {
                    t_w[k * VT_W + pl] = 0.0;
                }
            }
        }
        __syncwarp();
        tile_store_t(w, t_w0, tarr0, p0, np);
        if (use1) tile_store_t(w, t_w1, tarr1, p0, np);
        __syncwarp();
    }
    BwdSt b;
    b.cur = 0.0;
    double first = 0.0;
    for (int tl = ntile - 1; tl >= 0; --tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t_k, kap_pl, p0);
        tile_load_rows(w, t_e, el_pl, p0);
        tile_load_t(w, t_w0, tarr0, p0, np);
        if (use1) tile_load_t(w, t_w1, tarr1, p0, np);
        cp_async_wait_all();
        __syncwarp();
        if (compute) {
#pragma unroll 1
            for (int k = 31; k >= 0; --k) {
                const int p = p0 + k;
                if (p >= lo && p <= hi) {
                    const double kabs = fabs(t_k[k * VT_W + pl]);
                    const double wp = t_w[k * VT_W + pl];
                    if (p == hi)
                        bwd_init(b, wp, kabs);
                    else
                        t_w[k * VT_W + pl] = bwd_step(b, wp, kabs, t_e[k * VT_W + pl], wmax, c);
                    if (p == lo) first = b.cur;
                } else if (zero_tail && p > hi && p < n) {
                    t_w[k * VT_W + pl] = 0.0;
                }
            }
        }
        __syncwarp();
        tile_store_t(w, t_w0, tarr0, p0, np);
        if (use1) tile_store_t(w, t_w1, tarr1, p0, np);
        __syncwarp();
    }
    return first;
}

// STATE: stateful tick (ltpl_state.cuh) -- every path starts at its cut index + the vel_course rows (bf.trim), the planned
// velocity comes from bf.vel (pointed at vel_plan by the host), the follow-mode object distance from bf.obj_dist (k_ref)
template <bool STATE>
__global__ void __launch_bounds__(32)
k_vel_tiled(const LatDev lt, const LtplParams prm, const LtplDims dm, const LtplBuffers bf) {
    extern __shared__ __align__(16) unsigned char vt_smem[];
    double* tiles = reinterpret_cast<double*>(vt_smem);
    int* qs = reinterpret_cast<int*>(tiles + VT_NTILES * VT_TILE);
    int* ns = qs + VT_P;
    long long* in_base = reinterpret_cast<long long*>(ns + VT_P);
    long long* out_base = in_base + VT_P;
    __shared__ double s_axm[3 * LTPL_MAX_AXM];
    __shared__ double sp_wcapc[VT_P], sp_wnx[VT_P], sp_snx[VT_P];   // per path slot: final-pass parameters
    __shared__ int sp_idx_c[VT_P], sp_stop[VT_P], sp_mode[VT_P];
    stage_axm(prm, s_axm);
    const int lane = threadIdx.x;
    const int pl = lane % VT_P;          // path slot of this lane
    const int role = lane / VT_P;        // 0: complete profile, 1: control profile / searches; >= 2: tile moves only
    const bool compute = role < 2;
    const int B = dm.batch;
    const int nq = LTPL_NSLOT * B;
    const int n_follow = bf.queue_cnt[0], n_other = bf.queue_cnt[1];
    const int wf = (n_follow + VT_P - 1) / VT_P, wo = (n_other + VT_P - 1) / VT_P;
    const int wid = blockIdx.x;
    if (wid >= wf + wo) return;
    const bool follow_cls = wid < wf;
    const int t = follow_cls ? (wid * VT_P + pl) : ((wid - wf) * VT_P + pl);
    const bool live = follow_cls ? (t < n_follow) : (t < n_other);
    const int q = live ? bf.queue[(follow_cls ? 0 : nq) + t] : -1;
    const int b = live ? q % B : 0;
    int st = live ? bf.status[q] : 0;
    const int action = live ? bf.action_id[q] : LTPL_ACT_NONE;
    int n = live ? bf.path_len[q] : 0;
    int off_in = 0, pref = 0;
    if (STATE && live) {
        pref = bf.trim[4 * q + 3];
        off_in = bf.trim[4 * q + 2] + pref;
        n = max(n - off_in, 0);
    }
    const double vel_plan = live ? bf.vel[b] : 0.0;
    bool prefix = false;
    if (live && vel_plan > prm.vel_max + 0.1) {  // VPFB:106 brake prefix: the reference raises (see header)
        if (role == 0) atomicOr(&bf.sc_flags[b], LTPL_SC_BRAKE_PREFIX);
        prefix = true;
        n = 0;
    }
    if (role == 0) {
        qs[pl] = (n > 0) ? q : -1;
        ns[pl] = n;
        in_base[pl] = (long long)max(q, 0) * dm.p_max + off_in;
        out_base[pl] = (long long)max(q, 0) * dm.p_max + pref;
    }
    __syncwarp();
    int np = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) np = max(np, __shfl_xor_sync(LTPL_FULL, np, o));
    if (np == 0) return;
    const int ntile = (np + 31) >> 5;

    WarpCtx w;
    w.lane = lane;
    w.p_max = dm.p_max;
    w.ntc = nq + 64;
    w.col0 = wid * VT_P;
    w.qs = qs;
    w.ns = ns;
    w.in_base = in_base;
    w.out_base = out_base;
    const size_t pplane = (size_t)nq * dm.p_max;
    const double* x_pl = bf.path;
    const double* y_pl = bf.path + pplane;
    const double* k_pl = bf.path + 3 * pplane;
    const double* e_pl = bf.path + 4 * pplane;
    const size_t tsz = (size_t)dm.p_max * w.ntc;
    double* T_S = bf.vel_t;            // s
    double* T_B = T_S + tsz;           // ego brake profile
    double* T_C = T_B + tsz;           // control profile / reduced-horizon profile
    double* T_M = T_C + tsz;           // complete profile
    double* T_F = T_M + tsz;           // final w
    const int mycol = w.col0 + pl;
    const VelCfg c = make_velcfg(prm, s_axm);
    const double wmax = prm.vel_max * prm.vel_max;
    const bool red = (st & LTPL_ST_REDUCED_HORIZON) != 0;
    const bool any_red = __any_sync(LTPL_FULL, follow_cls && live && n > 0 && red);   // warp uniform
    bool vel_bound = true;

    double* t0 = tiles;
    double* t1 = tiles + VT_TILE;
    double* t2 = tiles + 2 * VT_TILE;
    double* t3 = tiles + 3 * VT_TILE;
    double* t4 = tiles + 4 * VT_TILE;
    double* t5 = tiles + 5 * VT_TILE;
    const int partner = (pl + VT_P) & 31;   // role-1 lane of this path

    // ------------------------------------------------------------------------------------------------------------------
    // pass A (forward).  role 0: s = [0, cumsum(el[:-1])] (OTH:743), ego brake profile (CVPF:152-165);
    //                    role 1: nearest path point to the object and to the ego position (OTH:774-782)
    // ------------------------------------------------------------------------------------------------------------------
    LTPL_PH_INIT
    const double ox = live ? bf.cobj[4 * b] : 0.0, oy = live ? bf.cobj[4 * b + 1] : 0.0;
    const double ov = live ? bf.cobj[4 * b + 2] : 0.0;
    const double epx = live ? bf.pos[2 * b] : 0.0, epy = live ? bf.pos[2 * b + 1] : 0.0;
    double acc_s = 0.0, spl_len = 0.0;
    double cur_b = 0.0, kb_prev = 0.0, eb_prev = 0.0, ego_stop_dist = 0.0;
    bool b_stopped = false, counting = true;
    double bv1 = LTPL_INF, bv2 = LTPL_INF;
    int nb1 = 0, nb2 = 0;
    for (int tl = 0; tl < ntile; ++tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t1, e_pl, p0);
        if (follow_cls) {
            tile_load_rows(w, t0, k_pl, p0);
            tile_load_rows(w, t2, x_pl, p0);
            tile_load_rows(w, t3, y_pl, p0);
        }
        cp_async_wait_all();
        __syncwarp();
        if (role == 0) {
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const int p = p0 + k;
                if (p < n) {
                    const double e = t1[k * VT_W + pl];
                    t4[k * VT_W + pl] = acc_s;   // s[p]
                    if (p < n - 1) spl_len = acc_s + e;
                    acc_s += e;
                    if (follow_cls) {
                        const double kabs = fabs(t0[k * VT_W + pl]);
                        if (p == 0) {
                            const double vs = fmax(vel_plan, 0.0);
                            cur_b = vs * vs;
                        } else if (!b_stopped) {
                            const double a = acc_brake(cur_b, kb_prev, c.ax_max, c.inv_ay, c.exp_, c.dm);
                            const double nx = fma(2.0 * a, eb_prev, cur_b);
                            if (nx < 0.0) {
                                b_stopped = true;
                                cur_b = 0.0;
                            } else {
                                cur_b = nx;
                            }
                        }
                        t5[k * VT_W + pl] = cur_b;
                        if (counting) {
                            if (cur_b > 0.01)
                                ego_stop_dist += e;
                            else
                                counting = false;
                        }
                        kb_prev = kabs;
                        eb_prev = e;
                    }
                }
            }
        } else if (role == 1 && follow_cls && !STATE) {
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const int p = p0 + k;
                if (p < n) {
                    const double xx = t2[k * VT_W + pl], yy = t3[k * VT_W + pl];
                    const double d1 = dist2_rn(xx, yy, ox, oy);
                    if (d1 < bv1) {
                        bv1 = d1;
                        nb1 = p;
                    }
                    const double d2 = dist2_rn(xx, yy, epx, epy);
                    if (d2 < bv2) {
                        bv2 = d2;
                        nb2 = p;
                    }
                }
            }
        }
        __syncwarp();
        tile_store_t(w, t4, T_S, p0, np);
        if (follow_cls) tile_store_t(w, t5, T_B, p0, np);
        __syncwarp();
    }
    LTPL_PH(0)
    __threadfence_block();
    const double* scol = T_S + mycol;
    // role exchange: both role lanes of a path continue with identical scalars
    spl_len = __shfl_sync(LTPL_FULL, spl_len, pl);
    ego_stop_dist = __shfl_sync(LTPL_FULL, ego_stop_dist, pl);
    nb1 = __shfl_sync(LTPL_FULL, nb1, partner);
    nb2 = __shfl_sync(LTPL_FULL, nb2, partner);

    // ------------------------------------------------------------------------------------------------------------------
    // per-path scalars of follow mode (CVPF:139-247)
    // ------------------------------------------------------------------------------------------------------------------
    int flags = 0;
    bool use_prof = false, has_ctrl = false;
    int idx_c = 0, stop_idx = 0;
    double vcs = 0.0, v_end_c = 0.0, v_control = 0.0;
    const double v_start_f = vel_plan;
    {
        double s_mine = 0.0;
        if (!STATE && follow_cls && compute && n > 0) {   // role 0: s of the object, role 1: s of the ego position
            const double* xr = x_pl + (size_t)q * dm.p_max;
            const double* yr = y_pl + (size_t)q * dm.p_max;
            const double* er = e_pl + (size_t)q * dm.p_max;
            s_mine = s_coord_from_nb(xr, yr, er, scol, w.ntc, n, role ? nb2 : nb1, role ? epx : ox, role ? epy : oy);
        }
        const double s_obj = __shfl_sync(LTPL_FULL, s_mine, pl);
        const double s_start = __shfl_sync(LTPL_FULL, s_mine, partner);
        if (follow_cls && compute && n > 0) {
            const double obj_dist = STATE ? bf.obj_dist[b] : (s_obj - s_start);   // OTH:784
            const double v_ego = bf.vel_est[b];
            const double control_d = prm.follow_c_p * prm.safety_d + lt.veh_length;
            const double safety_d = prm.safety_d + lt.veh_length;
            if ((obj_dist - safety_d) < 0) flags |= 1;
            const int ng = lt.n_glob - 1;
            const double* __restrict__ G = lt.glob_rl;
            const int start = bf.cobj_start[b];   // opponent on the closed global race line (k_plan, CVPF:166-179)
            double opp_stop_dist = 0.0;           // brake distance with ggv = [100, 14, 14] (CVPF:134, 185-199)
            {
                double v0 = fmin(ov, G[6 * start + 4]);
                if (v0 < 0.0) v0 = 0.0;
                double ww = v0 * v0;
                int id = 0;
                while (id < ng && ww > 0.01) {
                    int r = start + id;
                    if (r >= ng) r -= ng;
                    const double e = G[6 * r + 5];
                    opp_stop_dist += e;
                    ++id;
                    if (id <= ng - 1) {
                        const double a = acc_brake(ww, fabs(G[6 * r + 3]), 14.0, 1.0 / 14.0, c.exp_, c.dm);
                        const double nx = fma(2.0 * a, e, ww);
                        ww = (nx < 0.0) ? 0.0 : nx;
                    } else {
                        ww = 0.0;
                    }
                }
            }
            const double s_stop = obj_dist - safety_d + opp_stop_dist;   // CVPF:201-223
            stop_idx = min(first_ge_t(scol, w.ntc, n, s_stop), n - 1);
            const double s_last = scol[(size_t)(n - 1) * w.ntc];
            if (s_stop > s_last) {
                const double s_ends = opp_stop_dist - (s_stop - s_last);
                int idx = 0;
                double s_summed = 0.0;
                while (s_summed < s_ends && idx < ng) {
                    int r = start + idx;
                    if (r >= ng) r -= ng;
                    s_summed += G[6 * r + 5];
                    ++idx;
                }
                int r = start + idx;
                while (r >= ng) r -= ng;
                v_end_c = G[6 * r + 4];
            }
            if (prm.follow_control_type == 0) {   // CVPF:28-75
                v_control = ov - prm.follow_k_p * (control_d - obj_dist) + prm.follow_k_d * (ov - v_ego);
            } else {
                double arg = (control_d - obj_dist) * LTPL_PI / 2 * 1 / prm.follow_tan_w;
                arg = fmin(fmax(arg, -LTPL_PI / 2 + 1e-5), LTPL_PI / 2 - 1e-5);
                v_control = ov - tan(arg) * prm.follow_k_p + prm.follow_k_d * (ov - v_ego);
            }
            v_control = fmin(fmax(v_control, 0.0), prm.vel_max);
            if (ego_stop_dist < s_stop) {   // CVPF:247-292
                use_prof = true;
                const double* bcol = T_B + mycol;
                if (v_start_f > v_control && stop_idx >= 2) {
                    int first = first_le_t(bcol, w.ntc, n, v_control * v_control);
                    if (first >= n) first = 0;   // np.argmax of an all-False array
                    idx_c = min(first, stop_idx);
                    if (idx_c == 0) idx_c = stop_idx;
                    vcs = sqrt(bcol[(size_t)idx_c * w.ntc]);
                } else {
                    if (!(stop_idx >= 2)) flags |= 2;
                    idx_c = 0;
                    vcs = v_start_f;
                }
                has_ctrl = (stop_idx - idx_c) > 0;
            }
        }
    }

    LTPL_PH(1)
    // ------------------------------------------------------------------------------------------------------------------
    // follow: pass B / C.  role 0: complete profile on [0, n-1] -> T_M;  role 1: control profile on [idx_c, stop_idx]
    // (or the single value vcs^2 when stop_idx == idx_c) -> T_C                                        (CVPF:263-310)
    // ------------------------------------------------------------------------------------------------------------------
    if (follow_cls) {
        const double wmax_c = v_control * v_control;
        const double wcap_c = fmax(vcs, 0.0) * fmax(vcs, 0.0);
        const double wcap_m = fmax(v_start_f, 0.0) * fmax(v_start_f, 0.0);
        int lo, hi;
        double wcap, we, wmx;
        if (role == 0) {
            lo = 0;
            hi = n - 1;
            wcap = wcap_m;
            we = -1.0;
            wmx = wmax;
        } else {
            lo = idx_c;
            hi = (use_prof && has_ctrl) ? stop_idx : -1;
            wcap = wcap_c;
            we = fmax(v_end_c, 0.0) * fmax(v_end_c, 0.0);
            wmx = wmax_c;
        }
        const double w_first = profile_sweeps(w, tiles, k_pl, e_pl, np, pl, compute && n > 0, role, lo, hi, n, wcap, we,
                                              wmx, false, c, T_M, T_C, true);
        LTPL_PH(2)
        const double v0c = (role == 1 && use_prof && has_ctrl) ? sqrt(w_first) : vcs;
        const double v0c_r1 = __shfl_sync(LTPL_FULL, v0c, partner);
        // parameters of the intersection out = min(src, complete) for the fused final pass
        if (role == 0) {
            sp_idx_c[pl] = idx_c;
            sp_stop[pl] = stop_idx;
            sp_mode[pl] = (use_prof ? 1 : 0) | (has_ctrl ? 2 : 0);
            sp_wcapc[pl] = wcap_c;
        }
        __syncwarp();
        // a reduced-horizon follow path needs the intersection materialised (second profile + merge below)
        for (int tl = 0; tl < (any_red ? ntile : 0); ++tl) {
            const int p0 = tl << 5;
            tile_load_t(w, t2, T_C, p0, np);
            tile_load_t(w, t3, T_M, p0, np);
            tile_load_t(w, t5, T_B, p0, np);
            cp_async_wait_all();
            __syncwarp();
            if (role == 0) {
#pragma unroll 1
                for (int k = 0; k < 32; ++k) {
                    const int p = p0 + k;
                    if (p < n) {
                        double src = t5[k * VT_W + pl];   // ego brake profile
                        if (use_prof && p >= idx_c) {
                            if (p > stop_idx)
                                src = 0.0;
                            else
                                src = has_ctrl ? t2[k * VT_W + pl] : wcap_c;
                        }
                        t4[k * VT_W + pl] = fmin(src, t3[k * VT_W + pl]);
                    }
                }
            }
            __syncwarp();
            tile_store_t(w, t4, T_F, p0, np);
            __syncwarp();
        }
        if (use_prof) {
            if (has_ctrl && fabs(v0c_r1 - vcs) > 1.0) flags |= 2;
            const double prof0 = (idx_c == 0) ? v0c_r1 : fmax(v_start_f, 0.0);
            if (fabs(prof0 - v_start_f) > 1.0) flags |= 2;
        }
        if (flags & 1) st |= LTPL_ST_TOO_CLOSE;
        vel_bound = !(flags & 2);
    }

    LTPL_PH(3)
    // ------------------------------------------------------------------------------------------------------------------
    // all actions but follow, and follow with a reduced horizon: v_end rule + one profile on role 0 (OTH:834-923)
    // ------------------------------------------------------------------------------------------------------------------
    const bool need_single = live && n > 0 && (!follow_cls || red);
    if (__any_sync(LTPL_FULL, need_single)) {
        int hi = -1;
        double we = -1.0;
        if (need_single) {
            const int nn = bf.n_nodes[q];
            const int* nd = bf.nodes + ((size_t)q * dm.h_max + (nn - 1)) * 2;
            const int end_layer = nd[0], end_node = nd[1];
            int dn = end_node - lt.rl_idx[end_layer];
            if (dn < 0) dn = -dn;
            const double raceline_offset = dn * lt.lat_offset;   // quirk q3
            double v_end;
            int v_idx;
            if (red) {
                v_end = 0.0;
                // first i with cumsum(el[:-1])[i] >= spl_len - 5  <=>  s[i + 1] >= spl_len - 5   (OTH:851-859)
                int first = first_ge_t(scol + w.ntc, w.ntc, n - 1, spl_len - 5.0);
                if (first >= n - 1) first = 0;
                v_idx = first + 1;
                if (v_idx == 1 && n > 1) v_idx = n;
            } else {
                v_end = lt.vel_rl[end_layer];
                v_end -= fmin(v_end * lt.vel_decrease_lat * raceline_offset, v_end);
                v_idx = n;
            }
            if (v_idx > 1) {
                hi = v_idx - 1;
                we = fmax(v_end, 0.0) * fmax(v_end, 0.0);
            }
        }
        double* T_R = follow_cls ? T_C : T_F;   // follow: second profile into the (now free) control scratch
        const double wcap = fmax(vel_plan, 0.0) * fmax(vel_plan, 0.0);
        const double wf0 = profile_sweeps(w, tiles, k_pl, e_pl, np, pl, role == 0 && need_single, 0, 0, hi, n, wcap, we,
                                          wmax, true, c, T_R, T_R, false);
        if (need_single) {
            const double vf = (hi >= 0) ? sqrt(wf0) : 0.0;
            vel_bound = fabs(vf - vel_plan) < prm.v_max_offset;
        }
        vel_bound = __shfl_sync(LTPL_FULL, (int)vel_bound, pl) != 0;
        if (follow_cls) {
            // quirk q1 (OTH:923): row 5 decides column-wise -> the whole vx column comes from one of the two profiles
            bool take_second = false;
            if (need_single && n >= 6 && role == 0) {
                const double a5 = T_F[(size_t)5 * w.ntc + mycol], b5 = T_R[(size_t)5 * w.ntc + mycol];
                take_second = !(a5 < b5);
            }
            if (__any_sync(LTPL_FULL, take_second)) {
                for (int tl = 0; tl < ntile; ++tl) {
                    const int p0 = tl << 5;
                    tile_load_t(w, t0, T_F, p0, np);
                    tile_load_t(w, t1, T_R, p0, np);
                    cp_async_wait_all();
                    __syncwarp();
                    if (take_second)
                        for (int k = 0; k < 32; ++k) t0[k * VT_W + pl] = t1[k * VT_W + pl];
                    __syncwarp();
                    tile_store_t(w, t0, T_F, p0, np);
                    __syncwarp();
                }
            }
        }
    }

    LTPL_PH(4)
    // ------------------------------------------------------------------------------------------------------------------
    // pass D (backward): role 0: vx = sqrt(w), s;  role 1: ax = (w1 - w0) / (2 ds) with the standstill fix-up
    // (OTH:926-941); row-major output planes
    // ------------------------------------------------------------------------------------------------------------------
    // No recurrence here: every tile element is independent, so all 32 lanes work on the 32 x VT_P elements of a tile
    // (element e = it * 32 + lane -> point e / VT_P, path slot e % VT_P).  For follow paths without a reduced horizon
    // the intersection min(src, complete) (CVPF:297-310) is evaluated on the fly instead of in a sweep of its own.
    {
        const bool fused = follow_cls && !any_red;
        double* s_pl = bf.s_vx_ax;
        double* vx_pl = s_pl + pplane;
        double* ax_pl = vx_pl + pplane;
        if (lane < VT_P) {
            sp_wnx[lane] = 0.0;
            sp_snx[lane] = 0.0;
        }
        for (int tl = ntile - 1; tl >= 0; --tl) {
            const int p0 = tl << 5;
            if (fused) {
                tile_load_t(w, t0, T_C, p0, np);
                tile_load_t(w, t1, T_M, p0, np);
                tile_load_t(w, t2, T_B, p0, np);
            } else {
                tile_load_t(w, t1, T_F, p0, np);
            }
            tile_load_t(w, t3, T_S, p0, np);
            cp_async_wait_all();
            __syncwarp();
            if (fused) {
#pragma unroll
                for (int it = 0; it < VT_P; ++it) {
                    const int e = it * 32 + lane, k = e / VT_P, cc = e % VT_P, p = p0 + k;
                    if (p < ns[cc]) {
                        double src = t2[k * VT_W + cc];   // ego brake profile
                        const int mode = sp_mode[cc];
                        if ((mode & 1) && p >= sp_idx_c[cc]) {
                            if (p > sp_stop[cc])
                                src = 0.0;
                            else
                                src = (mode & 2) ? t0[k * VT_W + cc] : sp_wcapc[cc];
                        }
                        t1[k * VT_W + cc] = fmin(src, t1[k * VT_W + cc]);
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int it = 0; it < VT_P; ++it) {
                const int e = it * 32 + lane, k = e / VT_P, cc = e % VT_P, p = p0 + k;
                const int nn = ns[cc];
                if (p < nn) {
                    const double w0 = t1[k * VT_W + cc], s0 = t3[k * VT_W + cc];
                    double a = 0.0;
                    if (p < nn - 1) {
                        const double w1 = (k < 31) ? t1[(k + 1) * VT_W + cc] : sp_wnx[cc];
                        const double s1 = (k < 31) ? t3[(k + 1) * VT_W + cc] : sp_snx[cc];
                        a = (w1 - w0) / (2 * (s1 - s0));
                        if (w0 <= 1e-16 && fabs(a) <= 1e-8) a = -5.0;
                    }
                    t4[k * VT_W + cc] = sqrt(w0);
                    t5[k * VT_W + cc] = a;
                }
            }
            __syncwarp();
            if (lane < VT_P) {   // first row of this tile = successor of the last row of the next (lower) tile
                sp_wnx[lane] = t1[lane];
                sp_snx[lane] = t3[lane];
            }
            tile_store_rows(w, t3, s_pl, p0);
            tile_store_rows(w, t4, vx_pl, p0);
            tile_store_rows(w, t5, ax_pl, p0);
            __syncwarp();
        }
    }

    LTPL_PH(5)
    // acceptance (OTH:943-1025; no backup plan exists on the first tick)
    if (live && !prefix && n > 0 && role == 0) {
        if (!vel_bound) st |= LTPL_ST_VEL_BOUND_VIOL;
        // stateful tick: a backup plan exists (OTH:325-344), so a straight / follow profile that breaks the bound is
        // replaced by a brake profile on the OLD path (OTH:950-1006): flag here, k_backup plans it and clears the flag
        // (no backup plan exists after an invalid last solution, const_len == 0: the profile is kept, OTH:945-948)
        if (STATE && !vel_bound && (action == LTPL_ACT_FOLLOW || action == LTPL_ACT_STRAIGHT) && bf.const_len[b] != 0)
            atomicOr(&bf.sc_flags[b], LTPL_SC_STATE_FALLBACK | (6 << LTPL_SC_REASON_SHIFT));
        if (vel_bound || action == LTPL_ACT_FOLLOW || action == LTPL_ACT_STRAIGHT) {
            st |= LTPL_ST_TRAJ_VALID;
            bf.traj_len[q] = min(n, dm.n_export);
            bf.traj_id[q] = prm.traj_base_id + action;
            const int e = atomicAdd(&bf.queue_cnt[2], 1);
            bf.exp_q[e] = q;
            bf.traj_row[q] = e;
        }
        bf.status[q] = st;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_velprofile_tiled: stand-alone forward/backward solver over dense [n_paths][n_points] arrays (BASELINE config 5).
// Throughput regime (100 k paths): one warp = 32 paths, one lane = one path, tiles of 32 points x 32 paths.
//   sweep 1 (forward):  kappa, el tiles in (cp.async, transposing)  -> w = v^2 tile -> rows of `vx` (used as scratch)
//   sweep 2 (backward): kappa, el, w tiles in -> final w; vx = sqrt(w) and ax = (w1 - w0) / (2 el) in the same sweep
// HBM traffic: 5 reads + 3 writes of n_paths * n_points * 8 bytes (algorithmic minimum: 2 reads + 2 writes).
// ---------------------------------------------------------------------------------------------------------------------
#define VD_W 33
#ifndef VD_H
#define VD_H 16                            // points per tile (power of two <= 32): smaller tiles -> more resident warps
#endif
#define VD_TILE (VD_H * VD_W)
#define VD_SMEM_BYTES (3 * VD_TILE * 8)   // kappa, el, w tiles; vx / ax are produced in place

// one request covers 32 / VD_H path rows x VD_H consecutive points
__device__ __forceinline__ void vd_load(double* tile, const double* base, int path0, int n_paths, int n_points, int p0,
                                        int lane) {
    const int pt = lane % VD_H, sub = lane / VD_H;
#pragma unroll 4
    for (int r = sub; r < 32; r += 32 / VD_H) {
        if (path0 + r < n_paths && p0 + pt < n_points)
            cp_async8(&tile[pt * VD_W + r], base + (size_t)(path0 + r) * n_points + p0 + pt);
    }
}
__device__ __forceinline__ void vd_store(const double* tile, double* base, int path0, int n_paths, int n_points, int p0,
                                         int lane) {
    const int pt = lane % VD_H, sub = lane / VD_H;
#pragma unroll 4
    for (int r = sub; r < 32; r += 32 / VD_H) {
        if (path0 + r < n_paths && p0 + pt < n_points)
            base[(size_t)(path0 + r) * n_points + p0 + pt] = tile[pt * VD_W + r];
    }
}

__global__ void __launch_bounds__(32)
k_velprofile_tiled(const LtplParams prm, const LtplVelBatch vb) {
    extern __shared__ __align__(16) unsigned char vd_smem[];
    double* t_k = reinterpret_cast<double*>(vd_smem);
    double* t_e = t_k + VD_TILE;
    double* t_w = t_e + VD_TILE;
    double* t_v = t_w;   // sqrt(w) replaces w in place
    double* t_a = t_k;   // ax replaces kappa in place (kappa of the element is consumed before ax is written)
    __shared__ double s_axm[3 * LTPL_MAX_AXM];
    stage_axm(prm, s_axm);
    const int lane = threadIdx.x;
    const int path0 = blockIdx.x * 32;
    const int n = vb.n_points;
    const bool live = path0 + lane < vb.n_paths;
    const VelCfg c = make_velcfg(prm, s_axm);
    const double wmax = prm.vel_max * prm.vel_max;
    double vs = live ? vb.v_start[path0 + lane] : 0.0;
    double ve = live ? vb.v_end[path0 + lane] : 0.0;
    if (vs < 0.0) vs = 0.0;
    if (ve < 0.0) ve = 0.0;
    const int ntile = (n + VD_H - 1) / VD_H;
    FwdSt f;
    f.cur = 0.0;
    f.hint = 0;
    for (int tl = 0; tl < ntile; ++tl) {
        const int p0 = tl * VD_H;
        vd_load(t_k, vb.kappa, path0, vb.n_paths, n, p0, lane);
        vd_load(t_e, vb.el, path0, vb.n_paths, n, p0, lane);
        cp_async_wait_all();
        __syncwarp();
        if (live) {
#pragma unroll 1
            for (int k = 0; k < VD_H; ++k) {
                const int p = p0 + k;
                if (p < n) {
                    const double kabs = fabs(t_k[k * VD_W + lane]);
                    const double e = t_e[k * VD_W + lane];
                    const double oraw = c.ay_max / kabs;
                    double v = (p == 0) ? fwd_init(f, oraw, kabs, e, vs * vs, wmax) : fwd_step(f, oraw, kabs, e, wmax, c);
                    if (p == n - 1 && v > ve * ve) v = ve * ve;
                    t_w[k * VD_W + lane] = v;
                }
            }
        }
        __syncwarp();
        vd_store(t_w, vb.vx, path0, vb.n_paths, n, p0, lane);
        __syncwarp();
    }
    BwdSt b;
    b.cur = 0.0;
    double w_next = 0.0;
    for (int tl = ntile - 1; tl >= 0; --tl) {
        const int p0 = tl * VD_H;
        vd_load(t_k, vb.kappa, path0, vb.n_paths, n, p0, lane);
        vd_load(t_e, vb.el, path0, vb.n_paths, n, p0, lane);
        vd_load(t_w, vb.vx, path0, vb.n_paths, n, p0, lane);
        cp_async_wait_all();
        __syncwarp();
        if (live) {
#pragma unroll 1
            for (int k = VD_H - 1; k >= 0; --k) {
                const int p = p0 + k;
                if (p < n) {
                    const double kabs = fabs(t_k[k * VD_W + lane]);
                    const double e = t_e[k * VD_W + lane];
                    double wv = t_w[k * VD_W + lane];
                    double a = 0.0;
                    if (p == n - 1) {
                        bwd_init(b, wv, kabs);
                    } else {
                        wv = bwd_step(b, wv, kabs, e, wmax, c);
                        a = (w_next - wv) / (2 * e);
                    }
                    t_v[k * VD_W + lane] = sqrt(wv);
                    t_a[k * VD_W + lane] = a;
                    w_next = wv;
                }
            }
        }
        __syncwarp();
        vd_store(t_v, vb.vx, path0, vb.n_paths, n, p0, lane);
        vd_store(t_a, vb.ax, path0, vb.n_paths, n, p0, lane);
        __syncwarp();
    }
}
