// ltpl_vel_split.cuh -- k_vel_sweeps + k_vel_out: the velocity stage of ltpl_vel_tiled.cuh with the INDEPENDENT
// recurrences of a follow path running in different CTAs at the same time.
//
// k_vel_tiled is bound by the dependent chain of one warp: a follow path costs pass A (s, ego brake profile, n steps) +
// forward sweep (n) + backward sweep (n) + output.  Only the control profile needs the scalars of pass A; the complete
// profile (CVPF:297-310) and the reduced-horizon profile (OTH:834-923) do not.  k_vel_sweeps therefore gives every group
// of VT_P queued follow paths THREE one-warp CTAs that run concurrently on different SM sub-partitions:
//     type 0: pass A -> follow scalars -> control profile (forward + backward over [idx_c, stop_idx] only)   -> T_C
//     type 1: complete profile (forward + backward over the whole path)                                       -> T_M
//     type 2: (only when a path of the group has a reduced horizon) arc length, v_end rule, second profile    -> T_F
// and every group of other paths one CTA (type 3: arc length, v_end rule, one profile -> T_F).  The chain per warp drops
// from ~3 n to ~2 n steps.  k_vel_out (next kernel = the synchronisation point) is element-wise: intersection of the
// follow profiles (CVPF:297-310), quirk q1 (OTH:923), vx = sqrt(w), ax (OTH:926-941), acceptance (OTH:943-1025) and the
// compact export list.  Per-path scalars travel from k_vel_sweeps to k_vel_out in 8 doubles of `vel_scratch`.
#pragma once
#include "ltpl_vel_tiled.cuh"

#define VS_NTILES 4                      // k_vel_sweeps: kappa, el, w, curvature limit (pass A: kappa, el, x -> s, y -> brake)
#define VS_SMEM_BYTES (VS_NTILES * VT_TILE * 8 + 2 * VT_P * 4 + 2 * VT_P * 8)
#define VO_NTILES 7                      // k_vel_out
#define VO_SMEM_BYTES (VO_NTILES * VT_TILE * 8 + 2 * VT_P * 4 + 2 * VT_P * 8)
#ifndef VS_MINB
#define VS_MINB 20                      // resident one-warp CTAs per SM the register allocation of k_vel_sweeps is held to
#endif
#define VS_PARM 8   // doubles per path in vel_scratch: idx_c, stop_idx, mode, wcap_c, too_close, vel_bound, vel_bound2, -

struct VelGroup {
    int lane, pl, role;
    bool compute, follow_cls, live, prefix, red, any_red;
    int B, nq, g, q, b, st, action, n, np, ntile;
    double vel_plan;
};

// common prologue of both kernels: which VT_P queued paths this warp owns (group g of class follow / other)
__device__ __forceinline__ bool vel_group_init(VelGroup& G, const LtplParams& prm, const LtplDims& dm,
                                               const LtplBuffers& bf, int g, bool follow_cls, int idx_in_class, int* qs,
                                               int* ns, bool flag_prefix) {
    G.lane = threadIdx.x;
    G.pl = G.lane % VT_P;
    G.role = G.lane / VT_P;
    G.compute = G.role < 2;
    G.B = dm.batch;
    G.nq = LTPL_NSLOT * G.B;
    G.g = g;
    G.follow_cls = follow_cls;
    const int cnt = bf.queue_cnt[follow_cls ? 0 : 1];
    const int t = idx_in_class * VT_P + G.pl;
    G.live = t < cnt;
    G.q = G.live ? bf.queue[(follow_cls ? 0 : G.nq) + t] : -1;
    G.b = G.live ? G.q % G.B : 0;
    G.st = G.live ? bf.status[G.q] : 0;
    G.action = G.live ? bf.action_id[G.q] : LTPL_ACT_NONE;
    G.n = G.live ? bf.path_len[G.q] : 0;
    G.vel_plan = G.live ? bf.vel[G.b] : 0.0;
    G.prefix = false;
    if (G.live && G.vel_plan > prm.vel_max + 0.1) {  // VPFB:106 brake prefix: the reference raises (see ltpl_vel.cuh)
        if (flag_prefix && G.role == 0) atomicOr(&bf.sc_flags[G.b], LTPL_SC_BRAKE_PREFIX);
        G.prefix = true;
        G.n = 0;
    }
    if (G.role == 0) {
        qs[G.pl] = (G.n > 0) ? G.q : -1;
        ns[G.pl] = G.n;
        long long* base = reinterpret_cast<long long*>(ns + VT_P);   // in_base | out_base (first ticks: q * p_max)
        base[G.pl] = base[VT_P + G.pl] = (long long)max(G.q, 0) * dm.p_max;
    }
    __syncwarp();
    int np = G.n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) np = max(np, __shfl_xor_sync(LTPL_FULL, np, o));
    G.np = np;
    G.ntile = (np + 31) >> 5;
    G.red = (G.st & LTPL_ST_REDUCED_HORIZON) != 0;
    G.any_red = __any_sync(LTPL_FULL, follow_cls && G.live && G.n > 0 && G.red);
    return np > 0;
}

// One forward + one backward sweep over the tiles, ONE profile per compute lane (profile_sweeps of ltpl_vel_tiled.cuh
// without the second role; four tiles: kappa, el, w, curvature limit).  Profile on the path's points [lo, hi] (hi < lo:
// none): start cap wcap, end clamp we (< 0: none), speed limit wmax, result into the transposed array tarr; zero_tail:
// zeros behind hi up to n.  Returns w[lo] after the backward sweep.
__device__ __forceinline__ double profile_sweeps1(const WarpCtx& w, double* tiles, const double* kap_pl,
                                                  const double* el_pl, int np, int pl, bool compute, int lo, int hi, int n,
                                                  double wcap, double we, double wmax, bool zero_tail, const VelCfg& c,
                                                  double* tarr) {
    double* t_k = tiles;
    double* t_e = tiles + VT_TILE;
    double* t_w = tiles + 2 * VT_TILE;
    double* t_o = tiles + 3 * VT_TILE;
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
#pragma unroll
        for (int it = 0; it < VT_P; ++it) {   // curvature speed limit of the whole tile on all 32 lanes
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
        tile_store_t(w, t_w, tarr, p0, np);
        __syncwarp();
    }
    BwdSt b;
    b.cur = 0.0;
    double first = 0.0;
    for (int tl = ntile - 1; tl >= 0; --tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t_k, kap_pl, p0);
        tile_load_rows(w, t_e, el_pl, p0);
        tile_load_t(w, t_w, tarr, p0, np);
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
        tile_store_t(w, t_w, tarr, p0, np);
        __syncwarp();
    }
    return first;
}

// arc length of the path without its last element and the first index i with s[i + 1] >= spl_len - 5 (OTH:851-859),
// from the el tiles alone (role-0 lanes; used where the transposed s column is not available)
__device__ __forceinline__ void red_cut_index(const WarpCtx& w, double* tile, const double* e_pl, const VelGroup& G,
                                              double* spl_len_out, int* first_out) {
    double total = 0.0;
    for (int tl = 0; tl < G.ntile; ++tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, tile, e_pl, p0);
        cp_async_wait_all();
        __syncwarp();
        if (G.role == 0) {
#pragma unroll 1
            for (int k = 0; k < 32; ++k)
                if (p0 + k < G.n - 1) total += tile[k * VT_W + G.pl];
        }
        __syncwarp();
    }
    double acc = 0.0;
    int first = -1;
    for (int tl = 0; tl < G.ntile; ++tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, tile, e_pl, p0);
        cp_async_wait_all();
        __syncwarp();
        if (G.role == 0) {
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const int p = p0 + k;
                if (p < G.n - 1) {
                    acc += tile[k * VT_W + G.pl];   // == s[p + 1]
                    if (first < 0 && acc >= total - 5.0) first = p;
                }
            }
        }
        __syncwarp();
    }
    *spl_len_out = total;
    *first_out = first;
}

// v_end rule + range of the single profile of a path (OTH:834-865): hi = last point of the profile (-1: none), we = v_end^2
__device__ __forceinline__ void single_profile_range(const LatDev& lt, const LtplDims& dm, const LtplBuffers& bf,
                                                     const VelGroup& G, int first_red, int* hi_out, double* we_out) {
    const int nn = bf.n_nodes[G.q];
    const int* nd = bf.nodes + ((size_t)G.q * dm.h_max + (nn - 1)) * 2;
    const int end_layer = nd[0], end_node = nd[1];
    int dn = end_node - lt.rl_idx[end_layer];
    if (dn < 0) dn = -dn;
    const double raceline_offset = dn * lt.lat_offset;   // quirk q3
    double v_end;
    int v_idx;
    if (G.red) {
        v_end = 0.0;
        int first = first_red;
        if (first < 0 || first >= G.n - 1) first = 0;   // np.argmin of an all-False array
        v_idx = first + 1;
        if (v_idx == 1 && G.n > 1) v_idx = G.n;
    } else {
        v_end = lt.vel_rl[end_layer];
        v_end -= fmin(v_end * lt.vel_decrease_lat * raceline_offset, v_end);
        v_idx = G.n;
    }
    *hi_out = -1;
    *we_out = -1.0;
    if (v_idx > 1) {
        *hi_out = v_idx - 1;
        *we_out = fmax(v_end, 0.0) * fmax(v_end, 0.0);
    }
}

__global__ void __launch_bounds__(32, VS_MINB)
k_vel_sweeps(const LatDev lt, const LtplParams prm, const LtplDims dm, const LtplBuffers bf) {
    extern __shared__ __align__(16) unsigned char vt_smem[];
    double* tiles = reinterpret_cast<double*>(vt_smem);
    int* qs = reinterpret_cast<int*>(tiles + VS_NTILES * VT_TILE);
    int* ns = qs + VT_P;
    __shared__ double s_axm[3 * LTPL_MAX_AXM];
    const int n_follow = bf.queue_cnt[0], n_other = bf.queue_cnt[1];
    const int wf = (n_follow + VT_P - 1) / VT_P, wo = (n_other + VT_P - 1) / VT_P;
    const int bid = blockIdx.x;
    int type, g, idx;
    if (bid < 3 * wf) {          // heavy types first within a follow group: complete profile, control chain, reduced
        idx = bid / 3;
        g = idx;
        type = bid % 3;
        type = (type == 0) ? 1 : ((type == 1) ? 0 : 2);
    } else if (bid < 3 * wf + wo) {
        idx = bid - 3 * wf;
        g = wf + idx;
        type = 3;
    } else {
        return;
    }
    stage_axm(prm, s_axm);
    VelGroup G;
    if (!vel_group_init(G, prm, dm, bf, g, type != 3, idx, qs, ns, type == 0 || type == 3)) return;
    if (type == 2 && !G.any_red) return;
    const int lane = G.lane, pl = G.pl, role = G.role, n = G.n, q = G.q, b = G.b, np = G.np, ntile = G.ntile;
    const bool live = G.live, red = G.red;
    const double vel_plan = G.vel_plan;

    WarpCtx w;
    w.lane = lane;
    w.p_max = dm.p_max;
    w.ntc = G.nq + 64;
    w.col0 = g * VT_P;
    w.qs = qs;
    w.ns = ns;
    w.in_base = reinterpret_cast<const long long*>(ns + VT_P);
    w.out_base = w.in_base + VT_P;
    const size_t pplane = (size_t)G.nq * dm.p_max;
    const double* x_pl = bf.path;
    const double* y_pl = bf.path + pplane;
    const double* k_pl = bf.path + 3 * pplane;
    const double* e_pl = bf.path + 4 * pplane;
    const size_t tsz = (size_t)dm.p_max * w.ntc;
    double* T_S = bf.vel_t;            // s
    double* T_B = T_S + tsz;           // ego brake profile
    double* T_C = T_B + tsz;           // control profile
    double* T_M = T_C + tsz;           // complete profile
    double* T_F = T_M + tsz;           // single profile (other paths) / reduced-horizon profile (follow paths)
    const int mycol = w.col0 + pl;
    const VelCfg c = make_velcfg(prm, s_axm);
    const double wmax = prm.vel_max * prm.vel_max;
    double* parm = bf.vel_scratch + (size_t)(live ? q : 0) * VS_PARM;
    double* t0 = tiles;
    double* t1 = tiles + VT_TILE;
    double* t2 = tiles + 2 * VT_TILE;
    double* t3 = tiles + 3 * VT_TILE;
    const int partner = (pl + VT_P) & 31;   // role-1 lane of this path

    if (type == 1) {   // ---- complete profile on [0, n-1] (CVPF:297-306) ----
        const double wcap_m = fmax(vel_plan, 0.0) * fmax(vel_plan, 0.0);
        profile_sweeps1(w, tiles, k_pl, e_pl, np, pl, role == 0 && n > 0, 0, n - 1, n, wcap_m, -1.0, wmax, false, c, T_M);
        return;
    }

    if (type == 2 || type == 3) {   // ---- v_end rule + one profile (OTH:834-923) ----
        const bool need_single = live && n > 0 && (type == 3 || red);
        double spl_len = 0.0;
        int first_red = -1;
        if (type == 3) {   // arc length s (OTH:743) for the output pass
            double acc_s = 0.0;
            for (int tl = 0; tl < ntile; ++tl) {
                const int p0 = tl << 5;
                tile_load_rows(w, t1, e_pl, p0);
                cp_async_wait_all();
                __syncwarp();
                if (role == 0) {
#pragma unroll 1
                    for (int k = 0; k < 32; ++k) {
                        const int p = p0 + k;
                        if (p < n) {
                            t2[k * VT_W + pl] = acc_s;
                            if (p < n - 1) spl_len = acc_s + t1[k * VT_W + pl];
                            acc_s += t1[k * VT_W + pl];
                        }
                    }
                }
                __syncwarp();
                tile_store_t(w, t2, T_S, p0, np);
                __syncwarp();
            }
            __threadfence_block();
            if (need_single && red && role == 0) {
                // first i with cumsum(el[:-1])[i] >= spl_len - 5  <=>  s[i + 1] >= spl_len - 5   (OTH:851-859)
                first_red = first_ge_t(T_S + mycol + w.ntc, w.ntc, n - 1, spl_len - 5.0);
            }
        } else if (__any_sync(LTPL_FULL, need_single)) {
            red_cut_index(w, t1, e_pl, G, &spl_len, &first_red);
        }
        int hi = -1;
        double we = -1.0;
        if (need_single && role == 0) single_profile_range(lt, dm, bf, G, first_red, &hi, &we);
        const double wcap = fmax(vel_plan, 0.0) * fmax(vel_plan, 0.0);
        const double wf0 = profile_sweeps1(w, tiles, k_pl, e_pl, np, pl, role == 0 && need_single, 0, hi, n, wcap, we, wmax,
                                           true, c, T_F);
        if (need_single && role == 0) {
            const double vf = (hi >= 0) ? sqrt(wf0) : 0.0;
            const double vb = (fabs(vf - vel_plan) < prm.v_max_offset) ? 1.0 : 0.0;
            if (type == 3) {
                parm[2] = 0.0;
                parm[4] = 0.0;
                parm[5] = vb;
            } else {
                parm[6] = vb;
            }
        }
        return;
    }

    // ---- type 0: pass A (forward).  role 0: s = [0, cumsum(el[:-1])] (OTH:743), ego brake profile (CVPF:152-165);
    //                                 role 1: nearest path point to the object and to the ego position (OTH:774-782)
    const double ox = live ? bf.cobj[4 * b] : 0.0, oy = live ? bf.cobj[4 * b + 1] : 0.0;
    const double ov = live ? bf.cobj[4 * b + 2] : 0.0;
    const double epx = live ? bf.pos[2 * b] : 0.0, epy = live ? bf.pos[2 * b + 1] : 0.0;
    double acc_s = 0.0;
    double cur_b = 0.0, kb_prev = 0.0, eb_prev = 0.0, ego_stop_dist = 0.0;
    bool b_stopped = false, counting = true;
    double bv1 = LTPL_INF, bv2 = LTPL_INF;
    int nb1 = 0, nb2 = 0;
    for (int tl = 0; tl < ntile; ++tl) {
        const int p0 = tl << 5;
        tile_load_rows(w, t1, e_pl, p0);
        tile_load_rows(w, t0, k_pl, p0);
        tile_load_rows(w, t2, x_pl, p0);
        tile_load_rows(w, t3, y_pl, p0);
        cp_async_wait_all();
        __syncwarp();
        if (role == 1) {   // first the reader of the x / y tiles ...
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
        if (role == 0) {   // ... then s and the ego brake profile overwrite them
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
                const int p = p0 + k;
                if (p < n) {
                    const double e = t1[k * VT_W + pl];
                    t2[k * VT_W + pl] = acc_s;   // s[p]
                    acc_s += e;
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
                    t3[k * VT_W + pl] = cur_b;
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
        __syncwarp();
        tile_store_t(w, t2, T_S, p0, np);
        tile_store_t(w, t3, T_B, p0, np);
        __syncwarp();
    }
    __threadfence_block();
    const double* scol = T_S + mycol;
    ego_stop_dist = __shfl_sync(LTPL_FULL, ego_stop_dist, pl);
    nb1 = __shfl_sync(LTPL_FULL, nb1, partner);
    nb2 = __shfl_sync(LTPL_FULL, nb2, partner);

    // ---- per-path scalars of follow mode (CVPF:139-247) ----
    int flags = 0;
    bool use_prof = false, has_ctrl = false;
    int idx_c = 0, stop_idx = 0;
    double vcs = 0.0, v_end_c = 0.0, v_control = 0.0;
    const double v_start_f = vel_plan;
    {
        double s_mine = 0.0;
        if (G.compute && n > 0) {   // role 0: s of the object, role 1: s of the ego position
            const double* xr = x_pl + (size_t)q * dm.p_max;
            const double* yr = y_pl + (size_t)q * dm.p_max;
            const double* er = e_pl + (size_t)q * dm.p_max;
            s_mine = s_coord_from_nb(xr, yr, er, scol, w.ntc, n, role ? nb2 : nb1, role ? epx : ox, role ? epy : oy);
        }
        const double s_obj = __shfl_sync(LTPL_FULL, s_mine, pl);
        const double s_start = __shfl_sync(LTPL_FULL, s_mine, partner);
        if (G.compute && n > 0) {
            const double obj_dist = s_obj - s_start;   // OTH:784
            const double v_ego = bf.vel_est[b];
            const double control_d = prm.follow_c_p * prm.safety_d + lt.veh_length;
            const double safety_d = prm.safety_d + lt.veh_length;
            if ((obj_dist - safety_d) < 0) flags |= 1;
            const int ng = lt.n_glob - 1;
            const double* __restrict__ Gr = lt.glob_rl;
            const int start = bf.cobj_start[b];   // opponent on the closed global race line (k_plan, CVPF:166-179)
            double opp_stop_dist = 0.0;           // brake distance with ggv = [100, 14, 14] (CVPF:134, 185-199)
            {
                double v0 = fmin(ov, Gr[6 * start + 4]);
                if (v0 < 0.0) v0 = 0.0;
                double ww = v0 * v0;
                int id = 0;
                while (id < ng && ww > 0.01) {
                    int r = start + id;
                    if (r >= ng) r -= ng;
                    const double e = Gr[6 * r + 5];
                    opp_stop_dist += e;
                    ++id;
                    if (id <= ng - 1) {
                        const double a = acc_brake(ww, fabs(Gr[6 * r + 3]), 14.0, 1.0 / 14.0, c.exp_, c.dm);
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
                int idx2 = 0;
                double s_summed = 0.0;
                while (s_summed < s_ends && idx2 < ng) {
                    int r = start + idx2;
                    if (r >= ng) r -= ng;
                    s_summed += Gr[6 * r + 5];
                    ++idx2;
                }
                int r = start + idx2;
                while (r >= ng) r -= ng;
                v_end_c = Gr[6 * r + 4];
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

    // ---- control profile on [idx_c, stop_idx] (role-1 lanes; or the single value vcs^2 when stop_idx == idx_c) -> T_C ----
    const double wmax_c = v_control * v_control;
    const double wcap_c = fmax(vcs, 0.0) * fmax(vcs, 0.0);
    const int hi_c = (use_prof && has_ctrl) ? stop_idx : -1;
    const double we_c = fmax(v_end_c, 0.0) * fmax(v_end_c, 0.0);
    int hmax = (role == 1) ? hi_c : -1;   // sweep only as far as the longest control profile of the group reaches
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hmax = max(hmax, __shfl_xor_sync(LTPL_FULL, hmax, o));
    double w_first = 0.0;
    if (hmax >= 0) {
        const int np_c = min(np, hmax + 1);
        w_first = profile_sweeps1(w, tiles, k_pl, e_pl, np_c, pl, role == 1 && n > 0, idx_c, hi_c, min(n, np_c), wcap_c, we_c,
                                  wmax_c, false, c, T_C);
    }
    const double v0c = (role == 1 && use_prof && has_ctrl) ? sqrt(w_first) : vcs;
    const double v0c_r1 = __shfl_sync(LTPL_FULL, v0c, partner);
    if (use_prof) {
        if (has_ctrl && fabs(v0c_r1 - vcs) > 1.0) flags |= 2;
        const double prof0 = (idx_c == 0) ? v0c_r1 : fmax(v_start_f, 0.0);
        if (fabs(prof0 - v_start_f) > 1.0) flags |= 2;
    }
    if (live && n > 0 && role == 0) {
        parm[0] = (double)idx_c;
        parm[1] = (double)stop_idx;
        parm[2] = (double)((use_prof ? 1 : 0) | (has_ctrl ? 2 : 0));
        parm[3] = wcap_c;
        parm[4] = (flags & 1) ? 1.0 : 0.0;
        parm[5] = (flags & 2) ? 0.0 : 1.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_vel_out: element-wise tail of the velocity stage (see the header of this file)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_vel_out(const LatDev lt, const LtplParams prm, const LtplDims dm, const LtplBuffers bf) {
    extern __shared__ __align__(16) unsigned char vt_smem[];
    double* tiles = reinterpret_cast<double*>(vt_smem);
    int* qs = reinterpret_cast<int*>(tiles + VO_NTILES * VT_TILE);
    int* ns = qs + VT_P;
    __shared__ double sp_wcapc[VT_P], sp_wnx[VT_P], sp_snx[VT_P];
    __shared__ int sp_idx_c[VT_P], sp_stop[VT_P], sp_mode[VT_P];   // mode: 1 use_prof, 2 has_ctrl, 4 plain (read T_F)
    const int n_follow = bf.queue_cnt[0], n_other = bf.queue_cnt[1];
    const int wf = (n_follow + VT_P - 1) / VT_P, wo = (n_other + VT_P - 1) / VT_P;
    const int g = blockIdx.x;
    if (g >= wf + wo) return;
    const bool follow_cls = g < wf;
    VelGroup G;
    if (!vel_group_init(G, prm, dm, bf, g, follow_cls, follow_cls ? g : g - wf, qs, ns, false)) return;
    const int lane = G.lane, pl = G.pl, role = G.role, n = G.n, q = G.q, np = G.np, ntile = G.ntile;
    const bool live = G.live;

    WarpCtx w;
    w.lane = lane;
    w.p_max = dm.p_max;
    w.ntc = G.nq + 64;
    w.col0 = g * VT_P;
    w.qs = qs;
    w.ns = ns;
    w.in_base = reinterpret_cast<const long long*>(ns + VT_P);
    w.out_base = w.in_base + VT_P;
    const size_t pplane = (size_t)G.nq * dm.p_max;
    const size_t tsz = (size_t)dm.p_max * w.ntc;
    const double* T_S = bf.vel_t;
    const double* T_B = T_S + tsz;
    const double* T_C = T_B + tsz;
    const double* T_M = T_C + tsz;
    const double* T_F = T_M + tsz;
    const int mycol = w.col0 + pl;
    const double* parm = bf.vel_scratch + (size_t)(live ? q : 0) * VS_PARM;
    double* t0 = tiles;
    double* t1 = tiles + VT_TILE;
    double* t2 = tiles + 2 * VT_TILE;
    double* t3 = tiles + 3 * VT_TILE;
    double* t4 = tiles + 4 * VT_TILE;
    double* t5 = tiles + 5 * VT_TILE;
    double* t6 = tiles + 6 * VT_TILE;

    int st = G.st;
    bool vel_bound = true;
    int mode = 4, idx_c = 0, stop_idx = 0;
    double wcap_c = 0.0;
    if (live && n > 0 && role == 0) {
        if (follow_cls) {
            idx_c = (int)parm[0];
            stop_idx = (int)parm[1];
            mode = (int)parm[2];
            wcap_c = parm[3];
            if (parm[4] != 0.0) st |= LTPL_ST_TOO_CLOSE;
            vel_bound = parm[5] != 0.0;
            if (G.red) {
                vel_bound = parm[6] != 0.0;
                // quirk q1 (OTH:923): row 5 decides column-wise -> the whole vx column comes from one of the two profiles
                bool take_second = false;
                if (n >= 6) {
                    double src = T_B[(size_t)5 * w.ntc + mycol];
                    if ((mode & 1) && 5 >= idx_c) {
                        if (5 > stop_idx)
                            src = 0.0;
                        else
                            src = (mode & 2) ? T_C[(size_t)5 * w.ntc + mycol] : wcap_c;
                    }
                    const double a5 = fmin(src, T_M[(size_t)5 * w.ntc + mycol]);
                    const double b5 = T_F[(size_t)5 * w.ntc + mycol];
                    take_second = !(a5 < b5);
                }
                if (take_second) mode |= 4;
            }
        } else {
            vel_bound = parm[5] != 0.0;
        }
    }
    if (role == 0) {
        sp_idx_c[pl] = idx_c;
        sp_stop[pl] = stop_idx;
        sp_mode[pl] = mode;
        sp_wcapc[pl] = wcap_c;
    }
    if (lane < VT_P) {
        sp_wnx[lane] = 0.0;
        sp_snx[lane] = 0.0;
    }
    __syncwarp();
    const bool any_plain = __any_sync(LTPL_FULL, live && n > 0 && role == 0 && (mode & 4));

    // backward over the tiles: vx = sqrt(w), ax = (w1 - w0) / (2 ds) with the standstill fix-up (OTH:926-941); follow
    // paths evaluate the intersection min(src, complete) (CVPF:297-310) on the fly
    {
        double* s_pl = bf.s_vx_ax;
        double* vx_pl = s_pl + pplane;
        double* ax_pl = vx_pl + pplane;
        for (int tl = ntile - 1; tl >= 0; --tl) {
            const int p0 = tl << 5;
            if (follow_cls) {
                tile_load_t(w, t0, T_C, p0, np);
                tile_load_t(w, t1, T_M, p0, np);
                tile_load_t(w, t2, T_B, p0, np);
            }
            if (any_plain) tile_load_t(w, t6, T_F, p0, np);
            tile_load_t(w, t3, T_S, p0, np);
            cp_async_wait_all();
            __syncwarp();
#pragma unroll
            for (int it = 0; it < VT_P; ++it) {
                const int e = it * 32 + lane, k = e / VT_P, cc = e % VT_P, p = p0 + k;
                if (p < ns[cc]) {
                    const int md = sp_mode[cc];
                    double val;
                    if (md & 4) {
                        val = t6[k * VT_W + cc];
                    } else {
                        double src = t2[k * VT_W + cc];   // ego brake profile
                        if ((md & 1) && p >= sp_idx_c[cc]) {
                            if (p > sp_stop[cc])
                                src = 0.0;
                            else
                                src = (md & 2) ? t0[k * VT_W + cc] : sp_wcapc[cc];
                        }
                        val = fmin(src, t1[k * VT_W + cc]);
                    }
                    t1[k * VT_W + cc] = val;
                }
            }
            __syncwarp();
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

    // acceptance (OTH:943-1025; no backup plan exists on the first tick)
    if (live && !G.prefix && n > 0 && role == 0) {
        if (!vel_bound) st |= LTPL_ST_VEL_BOUND_VIOL;
        if (vel_bound || G.action == LTPL_ACT_FOLLOW || G.action == LTPL_ACT_STRAIGHT) {
            st |= LTPL_ST_TRAJ_VALID;
            bf.traj_len[q] = min(n, dm.n_export);
            bf.traj_id[q] = prm.traj_base_id + G.action;
            const int e = atomicAdd(&bf.queue_cnt[2], 1);
            bf.exp_q[e] = q;
            bf.traj_row[q] = e;
        }
        bf.status[q] = st;
    }
}
