// ltpl_vel_tiled.cuh -- k_vel_tiled: OTH.calc_vel_profile per action (OTH:688-1025) as TILE-STREAMED SIMT recurrences.
//
// Mapping: one warp = 32 queued paths of one class (follow / other), one thread = one path (the velocity solver is a
// serial recurrence per path).  A thread walking its own row of a [path][point] array touches a different cache line
// than its 31 neighbours on every access -- that made the first version of this kernel memory bound (ncu: long
// scoreboard stalls, ~1000 cycles per recurrence step).  Here every array is moved in TILES of 32 points x 32 paths:
//   * inputs (kappa, el, x, y; row-major per path) are copied global -> shared with cp.async, transposing on the fly
//     (one coalesced 256-byte request per path and tile; all requests of a tile are in flight together),
//   * the lanes then read tile[point][lane] (bank-conflict free) for 32 recurrence steps,
//   * intermediate profiles live in a TRANSPOSED global scratch [array][point][column] that only this warp touches
//     (coalesced rows, written from / read into tiles), results go back to the row-major planes through a tile.
// Passes over the tiles (follow class):  A forward  : s = cumsum(el), ego brake profile, nearest-point searches
//                                        B forward  : control profile + complete profile (two recurrences per step)
//                                        C backward : both backward passes, intersection with the brake profile
//                                        D backward : vx = sqrt(w), ax, row-major output
// (other class: A = cumsum only, B / C = one profile).  All arithmetic float64, w = v^2 domain (see ltpl_vel.cuh).
#pragma once
#include "ltpl_vel.cuh"

#define VT_W 33                      // padded tile row (doubles)
#define VT_TILE (32 * VT_W)          // doubles per tile
#define VT_NTILES 6                  // tiles per warp
#define VT_SMEM_BYTES (VT_NTILES * VT_TILE * 8 + 32 * 8)

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
    const int* qs;    // smem[32] path id per lane (-1 = idle lane)
    const int* ns;    // smem[32] points per lane (0 = idle lane)
};

// tile[k][r] = row_r[p0 + k] for the warp's 32 paths (row-major per-path array `plane`)
__device__ __forceinline__ void tile_load_rows(const WarpCtx& w, double* tile, const double* plane, int p0) {
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
        const int q = w.qs[r];
        if (p0 + w.lane < w.ns[r]) cp_async8(&tile[w.lane * VT_W + r], plane + (size_t)q * w.p_max + p0 + w.lane);
    }
}
// row_r[p0 + k] = tile[k][r]
__device__ __forceinline__ void tile_store_rows(const WarpCtx& w, const double* tile, double* plane, int p0) {
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
        const int q = w.qs[r];
        if (p0 + w.lane < w.ns[r]) plane[(size_t)q * w.p_max + p0 + w.lane] = tile[w.lane * VT_W + r];
    }
}
// transposed scratch <-> tile (rows of 32 consecutive columns)
__device__ __forceinline__ void tile_load_t(const WarpCtx& w, double* tile, const double* tarr, int p0, int np) {
#pragma unroll 4
    for (int k = 0; k < 32; ++k)
        if (p0 + k < np) cp_async8(&tile[k * VT_W + w.lane], tarr + (size_t)(p0 + k) * w.ntc + w.col0 + w.lane);
}
__device__ __forceinline__ void tile_store_t(const WarpCtx& w, const double* tile, double* tarr, int p0, int np) {
#pragma unroll 4
    for (int k = 0; k < 32; ++k)
        if (p0 + k < np) tarr[(size_t)(p0 + k) * w.ntc + w.col0 + w.lane] = tile[k * VT_W + w.lane];
}

// ---- recurrence state machines (same arithmetic as fb_profile_w / brake_profile_w in ltpl_vel.cuh) ----
struct FwdSt {
    double o_prev, cur, k_prev, e_prev;
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
    return o;
}
__device__ __forceinline__ double fwd_step(FwdSt& s, double oraw, double kabs, double e, double wmax, const VelCfg& c) {
    const double o_n = fmin(oraw, wmax);
    const bool rise = o_n > s.o_prev;
    if (!s.active && rise && !s.prev_rise) s.active = true;
    double nxt = o_n;
    if (s.active) {
        const double a = acc_forw(s.cur, s.k_prev, c, s.hint);
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

// get_s_coord.py:8-99 on an open polyline once the nearest index nb is known (x, y row-major; s transposed column)
__device__ __forceinline__ double s_coord_from_nb(const double* x, const double* y, const double* el, const double* scol,
                                                  int ntc, int n, int nb, double px, double py) {
    const int idx1 = max(nb - 1, 0), idx2 = min(nb + 1, n - 1);
    const double xn = x[nb], yn = y[nb];
    const double ang1 = fabs(angle3pt(xn, yn, px, py, x[idx1], y[idx1]));
    const double ang2 = fabs(angle3pt(xn, yn, px, py, x[idx2], y[idx2]));
    int ia, ib;
    if (ang1 > ang2) {
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

// generic single profile (others class; reduced-horizon second profile of follow): forward + backward over the tiles.
// range [0, hi] per lane (hi = -1: lane idle), w cap at start = wcap, end clamp we (< 0: none); result -> tarr_out
__device__ __forceinline__ double single_profile_passes(const WarpCtx& w, double* tiles, const double* kap_pl,
                                                        const double* el_pl, int np, int hi, double wcap, double we,
                                                        double wmax, const VelCfg& c, double* tarr_out) {
    double* t_k = tiles;
    double* t_e = tiles + VT_TILE;
    double* t_w = tiles + 2 * VT_TILE;
    const int ntile = (np + 31) >> 5;
    FwdSt f;
    f.cur = 0.0;
    for (int tl = 0; tl < ntile; ++tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t_k, kap_pl, p0);
        tile_load_rows(w, t_e, el_pl, p0);
        cp_async_wait_all();
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const int p = p0 + k;
            if (p <= hi) {
                const double kabs = fabs(t_k[k * VT_W + w.lane]);
                const double e = t_e[k * VT_W + w.lane];
                const double oraw = c.ay_max / kabs;
                double v = (p == 0) ? fwd_init(f, oraw, kabs, e, wcap, wmax) : fwd_step(f, oraw, kabs, e, wmax, c);
                if (p == hi && we >= 0.0 && v > we) v = we;
                t_w[k * VT_W + w.lane] = v;
            }
        }
        __syncwarp();
        tile_store_t(w, t_w, tarr_out, p0, np);
        __syncwarp();
    }
    BwdSt b;
    b.cur = 0.0;
    double first = 0.0;
    for (int tl = ntile - 1; tl >= 0; --tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t_k, kap_pl, p0);
        tile_load_rows(w, t_e, el_pl, p0);
        tile_load_t(w, t_w, tarr_out, p0, np);
        cp_async_wait_all();
        __syncwarp();
#pragma unroll 1
        for (int k = 31; k >= 0; --k) {
            const int p = p0 + k;
            if (p <= hi) {
                const double kabs = fabs(t_k[k * VT_W + w.lane]);
                const double wp = t_w[k * VT_W + w.lane];
                if (p == hi) {
                    bwd_init(b, wp, kabs);
                } else {
                    t_w[k * VT_W + w.lane] = bwd_step(b, wp, kabs, t_e[k * VT_W + w.lane], wmax, c);
                }
                if (p == 0) first = b.cur;
            } else if (p < w.ns[w.lane]) {
                t_w[k * VT_W + w.lane] = 0.0;   // zeros behind a shortened profile (OTH:900-903)
            }
        }
        __syncwarp();
        tile_store_t(w, t_w, tarr_out, p0, np);
        __syncwarp();
    }
    return first;
}

// final pass: vx = sqrt(w), ax = (w1 - w0) / (2 ds) with the standstill fix-up (OTH:926-941); row-major output planes
__device__ __forceinline__ void output_pass(const WarpCtx& w, double* tiles, const double* tarr_w, const double* tarr_s,
                                            int np, int n, double* s_pl, double* vx_pl, double* ax_pl) {
    double* t_w = tiles;
    double* t_s = tiles + VT_TILE;
    double* t_v = tiles + 2 * VT_TILE;
    double* t_a = tiles + 3 * VT_TILE;
    const int ntile = (np + 31) >> 5;
    double w_next = 0.0, s_next = 0.0;
    for (int tl = ntile - 1; tl >= 0; --tl) {
        const int p0 = tl << 5;
        tile_load_t(w, t_w, tarr_w, p0, np);
        tile_load_t(w, t_s, tarr_s, p0, np);
        cp_async_wait_all();
        __syncwarp();
#pragma unroll 1
        for (int k = 31; k >= 0; --k) {
            const int p = p0 + k;
            if (p < n) {
                const double w0 = t_w[k * VT_W + w.lane], s0 = t_s[k * VT_W + w.lane];
                double a = 0.0;
                if (p < n - 1) {
                    a = (w_next - w0) / (2 * (s_next - s0));
                    if (w0 <= 1e-16 && fabs(a) <= 1e-8) a = -5.0;
                }
                t_v[k * VT_W + w.lane] = sqrt(w0);
                t_a[k * VT_W + w.lane] = a;
                w_next = w0;
                s_next = s0;
            }
        }
        __syncwarp();
        tile_store_rows(w, t_s, s_pl, p0);
        tile_store_rows(w, t_v, vx_pl, p0);
        tile_store_rows(w, t_a, ax_pl, p0);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32)
k_vel_tiled(const LatDev lt, const LtplParams prm, const LtplDims dm, const LtplBuffers bf) {
    extern __shared__ __align__(16) unsigned char vt_smem[];
    double* tiles = reinterpret_cast<double*>(vt_smem);
    int* qs = reinterpret_cast<int*>(tiles + VT_NTILES * VT_TILE);
    int* ns = qs + 32;
    __shared__ double s_axm[3 * LTPL_MAX_AXM];
    stage_axm(prm, s_axm);
    const int lane = threadIdx.x;
    const int B = dm.batch;
    const int nq = LTPL_NSLOT * B;
    const int n_follow = bf.queue_cnt[0], n_other = bf.queue_cnt[1];
    const int wf = (n_follow + 31) >> 5, wo = (n_other + 31) >> 5;
    const int wid = blockIdx.x;
    if (wid >= wf + wo) return;
    const bool follow_cls = wid < wf;
    const int t = follow_cls ? (wid * 32 + lane) : ((wid - wf) * 32 + lane);
    const bool live = follow_cls ? (t < n_follow) : (t < n_other);
    const int q = live ? bf.queue[(follow_cls ? 0 : nq) + t] : -1;
    const int b = live ? q % B : 0;
    int st = live ? bf.status[q] : 0;
    const int action = live ? bf.action_id[q] : LTPL_ACT_NONE;
    int n = live ? bf.path_len[q] : 0;
    const double vel_plan = live ? bf.vel[b] : 0.0;
    bool prefix = false;
    if (live && vel_plan > prm.vel_max + 0.1) {  // VPFB:106 brake prefix: the reference raises (see header)
        atomicOr(&bf.sc_flags[b], LTPL_SC_BRAKE_PREFIX);
        prefix = true;
        n = 0;
    }
    qs[lane] = (n > 0) ? q : -1;
    ns[lane] = n;
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
    w.col0 = wid * 32;
    w.qs = qs;
    w.ns = ns;
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
    const int mycol = w.col0 + lane;
    const VelCfg c = make_velcfg(prm, s_axm);
    const double wmax = prm.vel_max * prm.vel_max;
    const bool red = (st & LTPL_ST_REDUCED_HORIZON) != 0;
    bool vel_bound = true;

    double* t0 = tiles;
    double* t1 = tiles + VT_TILE;
    double* t2 = tiles + 2 * VT_TILE;
    double* t3 = tiles + 3 * VT_TILE;
    double* t4 = tiles + 4 * VT_TILE;
    double* t5 = tiles + 5 * VT_TILE;

    // ------------------------------------------------------------------------------------------------------------------
    // pass A (forward): s = [0, cumsum(el[:-1])] (OTH:743); follow: ego brake profile (CVPF:152-165), nearest points
    // ------------------------------------------------------------------------------------------------------------------
    const double ox = live ? bf.cobj[4 * b] : 0.0, oy = live ? bf.cobj[4 * b + 1] : 0.0;
    const double ov = live ? bf.cobj[4 * b + 2] : 0.0;
    const double epx = live ? bf.pos[2 * b] : 0.0, epy = live ? bf.pos[2 * b + 1] : 0.0;
    LTPL_PH_INIT
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
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const int p = p0 + k;
            if (p < n) {
                const double e = t1[k * VT_W + lane];
                t4[k * VT_W + lane] = acc_s;   // s[p]
                if (p < n - 1) spl_len = acc_s + e;
                acc_s += e;
                if (follow_cls) {
                    const double kabs = fabs(t0[k * VT_W + lane]);
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
                    t5[k * VT_W + lane] = cur_b;
                    if (counting) {
                        if (cur_b > 0.01)
                            ego_stop_dist += e;
                        else
                            counting = false;
                    }
                    kb_prev = kabs;
                    eb_prev = e;
                    const double xx = t2[k * VT_W + lane], yy = t3[k * VT_W + lane];
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
    const double* scol = T_S + mycol;
    LTPL_PH(0)

    // ------------------------------------------------------------------------------------------------------------------
    // per-path scalars
    // ------------------------------------------------------------------------------------------------------------------
    int flags = 0;
    bool use_prof = false, has_ctrl = false;
    int idx_c = 0, stop_idx = 0;
    double vcs = 0.0, v_end_c = 0.0, v_control = 0.0, v_start_f = vel_plan;
    if (follow_cls && n > 0) {
        const double* xr = x_pl + (size_t)q * dm.p_max;
        const double* yr = y_pl + (size_t)q * dm.p_max;
        const double* er = e_pl + (size_t)q * dm.p_max;
        const double s_obj = s_coord_from_nb(xr, yr, er, scol, w.ntc, n, nb1, ox, oy);
        const double s_start = s_coord_from_nb(xr, yr, er, scol, w.ntc, n, nb2, epx, epy);
        const double obj_dist = s_obj - s_start;   // OTH:784
        const double v_ego = bf.vel_est[b];
        const double control_d = prm.follow_c_p * prm.safety_d + lt.veh_length;   // CVPF:139-142
        const double safety_d = prm.safety_d + lt.veh_length;
        if ((obj_dist - safety_d) < 0) flags |= 1;
        // opponent matched to the closed global race line (CVPF:166-179)
        const int ng = lt.n_glob - 1;
        const double* __restrict__ G = lt.glob_rl;
        int start;
        {
            double bv = LTPL_INF;
            int nb = 0;
#pragma unroll 4
            for (int i = 0; i < ng; ++i) {
                const double d = dist2_rn(G[6 * i + 1], G[6 * i + 2], ox, oy);
                if (d < bv) {
                    bv = d;
                    nb = i;
                }
            }
            const int i1 = (nb - 1 < 0) ? ng - 1 : nb - 1;
            const int i2 = (nb + 1 > ng - 1) ? 0 : nb + 1;
            const double a1 = fabs(angle3pt(G[6 * nb + 1], G[6 * nb + 2], ox, oy, G[6 * i1 + 1], G[6 * i1 + 2]));
            const double a2 = fabs(angle3pt(G[6 * nb + 1], G[6 * nb + 2], ox, oy, G[6 * i2 + 1], G[6 * i2 + 2]));
            start = (a1 >= a2) ? i1 : nb;
        }
        double opp_stop_dist = 0.0;   // brake distance with ggv = [100, 14, 14] (CVPF:134, 185-199)
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

    // ------------------------------------------------------------------------------------------------------------------
    // follow: pass B (forward) control + complete profile, pass C (backward) + intersection (CVPF:263-310)
    // ------------------------------------------------------------------------------------------------------------------
    double v_first = 0.0;
    LTPL_PH(1)
    if (follow_cls) {
        const double wmax_c = v_control * v_control;
        const double wcap_c = fmax(vcs, 0.0) * fmax(vcs, 0.0);
        const double wcap_m = fmax(v_start_f, 0.0) * fmax(v_start_f, 0.0);
        const double we_c = fmax(v_end_c, 0.0) * fmax(v_end_c, 0.0);
        FwdSt fc, fm;
        fc.cur = 0.0;
        fm.cur = 0.0;
        for (int tl = 0; tl < ntile; ++tl) {
            const int p0 = tl << 5;
            tile_load_rows(w, t0, k_pl, p0);
            tile_load_rows(w, t1, e_pl, p0);
            cp_async_wait_all();
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const int p = p0 + k;
                if (p < n) {
                    const double kabs = fabs(t0[k * VT_W + lane]);
                    const double e = t1[k * VT_W + lane];
                    const double oraw = c.ay_max / kabs;
                    t3[k * VT_W + lane] = (p == 0) ? fwd_init(fm, oraw, kabs, e, wcap_m, wmax)
                                                   : fwd_step(fm, oraw, kabs, e, wmax, c);
                    if (use_prof) {
                        double vc = 0.0;
                        if (has_ctrl && p >= idx_c && p <= stop_idx) {
                            vc = (p == idx_c) ? fwd_init(fc, oraw, kabs, e, wcap_c, wmax_c)
                                              : fwd_step(fc, oraw, kabs, e, wmax_c, c);
                            if (p == stop_idx && vc > we_c) vc = we_c;   // v_end clamp of the control profile
                        } else if (!has_ctrl && p == idx_c) {
                            vc = wcap_c;   // stop_idx == idx_c: vx_control = [vx_control_start]
                        }
                        t2[k * VT_W + lane] = vc;
                    }
                }
            }
            __syncwarp();
            tile_store_t(w, t2, T_C, p0, np);
            tile_store_t(w, t3, T_M, p0, np);
            __syncwarp();
        }
        LTPL_PH(2)
        BwdSt bc, bm;
        bc.cur = 0.0;
        bm.cur = 0.0;
        double v0c = vcs;
        for (int tl = ntile - 1; tl >= 0; --tl) {
            const int p0 = tl << 5;
            tile_load_rows(w, t0, k_pl, p0);
            tile_load_rows(w, t1, e_pl, p0);
            tile_load_t(w, t2, T_C, p0, np);
            tile_load_t(w, t3, T_M, p0, np);
            tile_load_t(w, t5, T_B, p0, np);
            cp_async_wait_all();
            __syncwarp();
#pragma unroll 1
            for (int k = 31; k >= 0; --k) {
                const int p = p0 + k;
                if (p < n) {
                    const double kabs = fabs(t0[k * VT_W + lane]);
                    const double e = t1[k * VT_W + lane];
                    double wm = t3[k * VT_W + lane];
                    if (p == n - 1)
                        bwd_init(bm, wm, kabs);
                    else
                        wm = bwd_step(bm, wm, kabs, e, wmax, c);
                    double src = t5[k * VT_W + lane];   // ego brake profile
                    if (use_prof) {
                        if (p >= idx_c) {
                            double wc = (p <= stop_idx) ? t2[k * VT_W + lane] : 0.0;
                            if (has_ctrl && p <= stop_idx) {
                                if (p == stop_idx)
                                    bwd_init(bc, wc, kabs);
                                else
                                    wc = bwd_step(bc, wc, kabs, e, wmax_c, c);
                                if (p == idx_c) v0c = sqrt(wc);
                            }
                            src = wc;
                        }
                    }
                    t4[k * VT_W + lane] = fmin(src, wm);
                }
            }
            __syncwarp();
            tile_store_t(w, t4, T_F, p0, np);
            __syncwarp();
        }
        if (use_prof) {
            if (has_ctrl && fabs(v0c - vcs) > 1.0) flags |= 2;
            const double prof0 = (idx_c == 0) ? v0c : fmax(v_start_f, 0.0);
            if (fabs(prof0 - v_start_f) > 1.0) flags |= 2;
        }
        if (flags & 1) st |= LTPL_ST_TOO_CLOSE;
        vel_bound = !(flags & 2);
    }

    // ------------------------------------------------------------------------------------------------------------------
    // all actions but follow, and follow with a reduced horizon: v_end rule + one profile (OTH:834-923)
    // ------------------------------------------------------------------------------------------------------------------
    LTPL_PH(3)
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
        const double wf0 = single_profile_passes(w, tiles, k_pl, e_pl, np, hi, wcap, we, wmax, c, T_R);
        if (need_single) {
            const double vf = (hi >= 0) ? sqrt(wf0) : 0.0;
            vel_bound = fabs(vf - vel_plan) < prm.v_max_offset;
        }
        if (follow_cls) {
            // quirk q1 (OTH:923): row 5 decides column-wise -> the whole vx column comes from one of the two profiles
            bool take_second = false;
            if (need_single && n >= 6) {
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
                        for (int k = 0; k < 32; ++k) t0[k * VT_W + lane] = t1[k * VT_W + lane];
                    __syncwarp();
                    tile_store_t(w, t0, T_F, p0, np);
                    __syncwarp();
                }
            }
        }
    }
    (void)v_first;
    LTPL_PH(4)

    // ------------------------------------------------------------------------------------------------------------------
    // pass D: vx, ax, s -> row-major planes; acceptance (OTH:943-1025, no backup plan on the first tick)
    // ------------------------------------------------------------------------------------------------------------------
    double* s_pl = bf.s_vx_ax;
    double* vx_pl = s_pl + pplane;
    double* ax_pl = vx_pl + pplane;
    output_pass(w, tiles, T_F, T_S, np, n, s_pl, vx_pl, ax_pl);
    LTPL_PH(5)

    if (live && !prefix && n > 0) {
        if (!vel_bound) st |= LTPL_ST_VEL_BOUND_VIOL;
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
