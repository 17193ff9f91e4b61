// ltpl_vel.cuh -- forward/backward ggv-limited velocity profiles.
//   device functions: acc_tire / acc_forw / acc_decel (tph.calc_ax_poss), brake_profile_w (tph.calc_vel_profile_brake),
//   fb_profile_w (tph.calc_vel_profile, closed=False, loc_gg mode), follow_profile
//   (helper_funcs/calc_vel_profile_follow.py:78-313)
//   kernels: k_vel (OTH.calc_vel_profile per action, OTH:688-1025), k_export (OTH:941 + LTPL:401-406),
//            k_velprofile_dense (stand-alone solver over dense arrays, BASELINE config 5)
//
// One THREAD per path: the solver is a serial recurrence over the points of one path (SURVEY hard part 4), so the
// kernel is bound by the latency of the dependent chain of one path, not by bandwidth.  Two algebraic rewrites keep
// that chain short while staying in float64:
//   * the recurrences are carried in w = v^2: v_next^2 = v^2 + 2 a(v^2) ds needs no sqrt and no division on the chain
//     (ay_used = v^2 / radius = w * |kappa|); only the machine limit of the forward pass needs v = sqrt(w).
//     vx = sqrt(w) and ax = (w1 - w0) / (2 ds) are evaluated afterwards, off the chain.  Differences to the
//     reference's v-domain arithmetic are O(1e-16) relative.
//   * "is index i the start of an acceleration phase" (tph scans the INITIAL profile for rising edges) only needs the
//     original values at i-1, i, i+1, which are still unmodified when the single forward / backward scan reaches i.
#pragma once
#include "ltpl_common.cuh"

struct VelCfg {
    double ax_max, ay_max, inv_ay;  // local gg * gg_scale (VPFB:213-214)
    double exp_, dm;                // friction-ellipse exponent, drag_coeff / m_veh
    const double* axm_v;
    const double* axm_a;
    const double* axm_s;
    int n_axm;
};

// The machine table is indexed with a per-lane velocity: read from the kernel-parameter (constant) bank that is a
// divergent constant load which the hardware serialises lane by lane (measured: ~2.8 k cycles per recurrence step).
// Every velocity kernel therefore first copies the three small tables to shared memory (stage_axm) and points the
// configuration at that copy.
__device__ __forceinline__ void stage_axm(const LtplParams& prm, double* s_axm /* [3 * LTPL_MAX_AXM] shared */) {
    for (int i = threadIdx.x; i < LTPL_MAX_AXM; i += blockDim.x) {
        s_axm[i] = prm.axm_v[i];
        s_axm[LTPL_MAX_AXM + i] = prm.axm_a[i];
        s_axm[2 * LTPL_MAX_AXM + i] = prm.axm_s[i];
    }
    __syncthreads();
}

__device__ __forceinline__ VelCfg make_velcfg(const LtplParams& prm, const double* s_axm) {
    VelCfg c;
    c.ax_max = prm.gg_ax * prm.gg_scale;
    c.ay_max = prm.gg_ay * prm.gg_scale;
    c.inv_ay = 1.0 / c.ay_max;
    c.exp_ = prm.dyn_model_exp;
    c.dm = prm.drag_coeff / prm.m_veh;
    c.axm_v = s_axm;
    c.axm_a = s_axm + LTPL_MAX_AXM;
    c.axm_s = s_axm + 2 * LTPL_MAX_AXM;
    c.n_axm = prm.n_axm;
    return c;
}

// available longitudinal tyre acceleration at w = v^2 on curvature |kappa| (friction ellipse with exponent exp)
// general friction-ellipse exponent: two pow() calls = ~1000 instructions; kept out of line so that the recurrence
// loops of the common exponent 1.0 (LTPL:190 default) stay small in the instruction cache
__device__ __noinline__ double acc_tire_pow(double ratio, double ax_max, double exp_) {
    const double radicand = 1.0 - pow(ratio, exp_);
    return (radicand > 0.0) ? ax_max * pow(radicand, 1.0 / exp_) : 0.0;
}
__device__ __forceinline__ double acc_tire(double w, double kabs, double ax_max, double inv_ay, double exp_) {
    const double ratio = w * kabs * inv_ay;  // ay_used / ay_max, ay_used = v^2 / radius
    if (exp_ == 1.0) {
        const double radicand = 1.0 - ratio;
        return (radicand > 0.0) ? ax_max * radicand : 0.0;
    }
    return acc_tire_pow(ratio, ax_max, exp_);
}

// np.interp on the machine table with a moving hint (v changes slowly along a path)
__device__ __forceinline__ double interp_hint(double v, const double* __restrict__ xp, const double* __restrict__ fp,
                                              const double* __restrict__ sp, int n, int& j) {
    if (v <= xp[0]) return fp[0];
    if (v >= xp[n - 1]) return fp[n - 1];
    while (j < n - 2 && v >= xp[j + 1]) ++j;
    while (j > 0 && v < xp[j]) --j;
    return fma(sp[j], v - xp[j], fp[j]);
}

// mode 'accel_forw': min(tyre, machine(v)) + drag,  drag = -v^2 drag_coeff / m
__device__ __forceinline__ double acc_forw(double w, double kabs, const VelCfg& c, int& hint) {
    double a = acc_tire(w, kabs, c.ax_max, c.inv_ay, c.exp_);
    const double axm = interp_hint(sqrt(w), c.axm_v, c.axm_a, c.axm_s, c.n_axm, hint);
    a = fmin(a, axm);
    return fma(-w, c.dm, a);
}

// mode 'decel_backw': tyre - drag
__device__ __forceinline__ double acc_backw(double w, double kabs, const VelCfg& c) {
    return fma(w, c.dm, acc_tire(w, kabs, c.ax_max, c.inv_ay, c.exp_));
}

// mode 'decel_forw' with ggv (ax_max, ay_max): -tyre + drag (both negative)
__device__ __forceinline__ double acc_brake(double w, double kabs, double ax_max, double inv_ay, double exp_, double dm) {
    return fma(-w, dm, -acc_tire(w, kabs, ax_max, inv_ay, exp_));
}

// tph.calc_vel_profile_brake in w: w[0] = v_start^2, forward integration with full braking, zeros after standstill.
// returns the number of leading entries with v > 0.1 (== "id_brake" of CVPF:161-163) and their summed element length.
__device__ __forceinline__ int brake_profile_w(const double* __restrict__ kap, const double* __restrict__ el, int n,
                                               double v_start, const VelCfg& c, double* w, double* stop_dist) {
    if (v_start < 0.0) v_start = 0.0;
    double cur = v_start * v_start;
    w[0] = cur;
    int i = 0, id_brake = 0;
    double dist = 0.0;
    bool counting = true;
    for (; i < n - 1; ++i) {
        const double e = el[i];
        if (counting) {
            if (cur > 0.01) {
                ++id_brake;
                dist += e;
            } else {
                counting = false;
            }
        }
        const double a = acc_brake(cur, fabs(kap[i]), c.ax_max, c.inv_ay, c.exp_, c.dm);
        const double nx = fma(2.0 * a, e, cur);
        if (nx < 0.0) break;
        cur = nx;
        w[i + 1] = cur;
    }
    if (i == n - 1) {  // ran to the end without standstill: the last entry still has to be counted
        if (counting && cur > 0.01) {
            ++id_brake;
            dist += el[n - 1];
        }
    } else {
        for (int k = i + 1; k < n; ++k) w[k] = 0.0;
    }
    *stop_dist = dist;
    return id_brake;
}

// tph.calc_vel_profile(closed=False, loc_gg mode) in w = v^2.  Returns w[0] after the backward pass.
__device__ __forceinline__ double fb_profile_w(const double* __restrict__ kap, const double* __restrict__ el, int n,
                                               double v_max, double v_start, bool has_end, double v_end,
                                               const VelCfg& c, double* w) {
    if (v_start < 0.0) v_start = 0.0;
    if (has_end && v_end < 0.0) v_end = 0.0;
    const double wmax = v_max * v_max;
    int hint = 0;
    // ---- forward (mode accel_forw) ----
    double k_i = fabs(kap[0]);
    double o_i = fmin(c.ay_max / k_i, wmax);  // (sqrt(ay * radius))^2, radius = 1 / |kappa| (inf for kappa == 0)
    o_i = fmin(o_i, v_start * v_start);
    double cur = o_i;
    w[0] = cur;
    bool prev_rise = false, active = false;
    double k_n = (n > 1) ? fabs(kap[1]) : 0.0;
    double e_i = (n > 1) ? el[0] : 0.0;
    for (int i = 0; i < n - 1; ++i) {
        // software pipelining: next iteration's operands are requested before this iteration's dependent math
        const double k_nn = (i + 2 < n) ? fabs(kap[i + 2]) : 0.0;
        const double e_n = (i + 1 < n - 1) ? el[i + 1] : 0.0;
        const double o_n = fmin(c.ay_max / k_n, wmax);
        const bool rise = o_n > o_i;
        if (!active && rise && !prev_rise) active = true;
        double nxt = o_n;
        if (active) {
            const double a = acc_forw(cur, k_i, c, hint);
            const double wn = fma(2.0 * a, e_i, cur);
            if (wn < o_n) nxt = wn;
            if (wn > wmax) active = false;
        }
        w[i + 1] = nxt;
        cur = nxt;
        prev_rise = rise;
        o_i = o_n;
        k_i = k_n;
        k_n = k_nn;
        e_i = e_n;
    }
    if (has_end) {
        const double we = v_end * v_end;
        if (cur > we) {
            cur = we;
            w[n - 1] = cur;
        }
    }
    // ---- backward (flipped arrays, mode decel_backw, one look-ahead correction) ----
    o_i = cur;
    prev_rise = false;
    active = false;
    double k_p = fabs(kap[n - 1]);
    double o_n = (n > 1) ? w[n - 2] : 0.0;
    double k_pn = (n > 1) ? fabs(kap[n - 2]) : 0.0;
    double e_pn = (n > 1) ? el[n - 2] : 0.0;
    for (int j = 0; j < n - 1; ++j) {
        const int pn = n - 2 - j;
        const double o_nn = (pn > 0) ? w[pn - 1] : 0.0;
        const double k_pnn = (pn > 0) ? fabs(kap[pn - 1]) : 0.0;
        const double e_pnn = (pn > 0) ? el[pn - 1] : 0.0;
        const bool rise = o_n > o_i;
        if (!active && rise && !prev_rise) active = true;
        double nxt = o_n;
        if (active) {
            const double a = acc_backw(cur, k_p, c);
            double wn = fma(2.0 * a, e_pn, cur);
            const double a2 = acc_backw(wn, k_pn, c);
            const double wt = fma(2.0 * a2, e_pn, cur);
            wn = fmin(wn, wt);
            if (wn < o_n) nxt = wn;
            if (wn > wmax) active = false;
            w[pn] = nxt;
        }
        cur = nxt;
        prev_rise = rise;
        o_i = o_n;
        o_n = o_nn;
        k_p = k_pn;
        k_pn = k_pnn;
        e_pn = e_pnn;
    }
    return cur;
}

// get_s_coord.py:8-99 on an OPEN polyline given as planes x[], y[] with s_array = np.cumsum(el) (serial, one thread)
__device__ __forceinline__ double s_coord_open_path(const double* __restrict__ x, const double* __restrict__ y,
                                                    const double* __restrict__ s, const double* __restrict__ el, int n,
                                                    double px, double py) {
    double bv = LTPL_INF;
    int nb = 0;
    for (int i = 0; i < n; ++i) {
        const double d = dist2_rn(x[i], y[i], px, py);
        if (d < bv) {
            bv = d;
            nb = i;
        }
    }
    const int idx1 = max(nb - 1, 0), idx2 = min(nb + 1, n - 1);
    const double ang1 = fabs(angle3pt(x[nb], y[nb], px, py, x[idx1], y[idx1]));
    const double ang2 = fabs(angle3pt(x[nb], y[nb], px, py, x[idx2], y[idx2]));
    int ia, ib;
    if (ang1 > ang2) {
        ia = idx1;
        ib = nb;
    } else {
        ia = nb;
        ib = idx2;
    }
    // s_array = cumsum(el); a leading 0 is inserted when s_array[0] > 0.05 (get_s_coord.py:67-68):
    // with insertion s_array'[i] = sum(el[:i]) = s[i]; without it s_array[i] = s[i] + el[i]
    const bool ins = el[0] > 0.05;
    const double sbase = ins ? s[ia] : __dadd_rn(s[ia], el[ia]);
    const double ax = x[ia], ay = y[ia], bx = x[ib] - ax, by = y[ib] - ay;
    const double t = __ddiv_rn(__dadd_rn(__dmul_rn(px - ax, bx), __dmul_rn(py - ay, by)), __dadd_rn(sq_rn(bx), sq_rn(by)));
    const double sx = __dadd_rn(ax, __dmul_rn(t, bx)), sy = __dadd_rn(ay, __dmul_rn(t, by));
    const double ds = sqrt(__dadd_rn(sq_rn(ax - sx), sq_rn(ay - sy)));
    return __dadd_rn(sbase, ds);
}

// calc_vel_profile_follow (CVPF:78-313) in w = v^2.  kap / el / s have n entries (el[n-1] == 0).
// wb, prof, compl_: n-entry scratch rows.  returns flags: bit0 too_close, bit1 vel_bound violated; the result
// (np.minimum(prof, compl), squared) is written to `out`.
__device__ __forceinline__ int follow_profile(const LatDev& lt, const LtplParams& prm, const VelCfg& c,
                                              const double* __restrict__ kap, const double* __restrict__ el,
                                              const double* __restrict__ s, int n, double v_start, double v_ego,
                                              double v_obj, double obj_dist, int glob_start, double* wb,
                                              double* prof, double* compl_, double* out) {
    int flags = 0;
    const double v_max = prm.vel_max;
    const double control_d = prm.follow_c_p * prm.safety_d + lt.veh_length;
    const double safety_d = prm.safety_d + lt.veh_length;
    if ((obj_dist - safety_d) < 0) flags |= 1;

    // ego brake profile on the local path (CVPF:152-165)
    LTPL_PH_INIT
    double ego_stop_dist;
    brake_profile_w(kap, el, n, v_start, c, wb, &ego_stop_dist);
    LTPL_PH(2)

    // opponent matched to the (closed) global race line, rolled to start at its position (CVPF:166-179):
    // `start` = closest_indexes[0], computed warp-parallel in k_plan (LtplBuffers.cobj_start)
    const int ng = lt.n_glob - 1;
    const double* __restrict__ G = lt.glob_rl;
    const int start = glob_start;
    LTPL_PH(3)
    // opponent brake profile with ggv = [100, 14, 14] (CVPF:134, 185-199): only the stop distance is needed
    double opp_stop_dist = 0.0;
    {
        double v0 = fmin(v_obj, G[6 * start + 4]);
        if (v0 < 0.0) v0 = 0.0;
        double w = v0 * v0;
        int id = 0;
        while (id < ng && w > 0.01) {
            int r = start + id;
            if (r >= ng) r -= ng;
            const double e = G[6 * r + 5];
            opp_stop_dist += e;
            ++id;
            if (id <= ng - 1) {
                const double a = acc_brake(w, fabs(G[6 * r + 3]), 14.0, 1.0 / 14.0, c.exp_, c.dm);
                const double nx = fma(2.0 * a, e, w);
                w = (nx < 0.0) ? 0.0 : nx;
            } else {
                w = 0.0;
            }
        }
    }

    // characteristic positions (CVPF:201-223)
    int stop_idx = 0;
    const double s_stop = obj_dist - safety_d + opp_stop_dist;
    while (stop_idx < n - 1 && s[stop_idx] < s_stop) ++stop_idx;
    double v_end = 0.0;
    if (s_stop > s[n - 1]) {
        const double s_ends = opp_stop_dist - (s_stop - s[n - 1]);
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
        v_end = G[6 * r + 4];
    }

    // control velocity (CVPF:28-75, 232-239)
    double v_control;
    if (prm.follow_control_type == 0) {
        v_control = v_obj - prm.follow_k_p * (control_d - obj_dist) + prm.follow_k_d * (v_obj - v_ego);
    } else {
        double arg = (control_d - obj_dist) * LTPL_PI / 2 * 1 / prm.follow_tan_w;
        arg = fmin(fmax(arg, -LTPL_PI / 2 + 1e-5), LTPL_PI / 2 - 1e-5);
        v_control = v_obj - tan(arg) * prm.follow_k_p + prm.follow_k_d * (v_obj - v_ego);
    }
    v_control = fmin(fmax(v_control, 0.0), v_max);

    LTPL_PH(4)
    const double* src = wb;
    if (ego_stop_dist < s_stop) {
        int idx_c;
        double vcs;
        if (v_start > v_control && stop_idx >= 2) {
            const double wc = v_control * v_control;
            int first = 0;
            for (int i = 0; i < n; ++i)
                if (wb[i] <= wc) {
                    first = i;
                    break;
                }
            idx_c = min(first, stop_idx);
            if (idx_c == 0) idx_c = stop_idx;
            vcs = sqrt(wb[idx_c]);
        } else {
            if (!(stop_idx >= 2)) flags |= 2;
            idx_c = 0;
            vcs = v_start;
        }
        for (int i = 0; i < idx_c; ++i) prof[i] = wb[i];
        double v0c = vcs;
        if (stop_idx - idx_c > 0) {
            v0c = sqrt(fb_profile_w(kap + idx_c, el + idx_c, stop_idx - idx_c + 1, v_control, vcs, true, v_end, c,
                                    prof + idx_c));
            if (fabs(v0c - vcs) > 1.0) flags |= 2;
        } else {
            prof[idx_c] = vcs * vcs;
        }
        for (int i = stop_idx + 1; i < n; ++i) prof[i] = 0.0;
        const double prof0 = (idx_c == 0) ? v0c : fmax(v_start, 0.0);
        if (fabs(prof0 - v_start) > 1.0) flags |= 2;
        src = prof;
    }
    LTPL_PH(5)
    // complete (unconstrained) profile and intersection (CVPF:296-310)
    fb_profile_w(kap, el, n, v_max, v_start, false, 0.0, c, compl_);
    LTPL_PH(6)
    for (int i = 0; i < n; ++i) out[i] = fmin(src[i], compl_[i]);
    LTPL_PH(7)
    return flags;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_vel: OTH.get_ref_idx (never planned before, OTH:590-599) + OTH.calc_vel_profile per action (OTH:688-1025).
// Work items come from two dense queues filled by k_path (class 0: follow, class 1: straight / left / right), so that
// every warp is full and runs one code path.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LTPL_VEL_BLOCK)
k_vel(const LatDev lt, const LtplParams prm, const LtplDims dm, const LtplBuffers bf) {
    __shared__ double s_axm[3 * LTPL_MAX_AXM];
    stage_axm(prm, s_axm);
    const int B = dm.batch;
    const int nq = LTPL_NSLOT * B;
    // LTPL_VEL_LANES work items per warp (remaining lanes idle): the kernel is bound by the memory / dependent-issue
    // latency of one path and there are only ~13 k paths per 10 k-scenario batch, so spreading them over more warps
    // buys latency overlap per SM at the price of issue slots that would be idle anyway.
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    if ((gt & 31) >= LTPL_VEL_LANES) return;
    const int t = (gt >> 5) * LTPL_VEL_LANES + (gt & 31);
    const int n_follow = bf.queue_cnt[0], n_other = bf.queue_cnt[1];
    const int n_follow_pad = (n_follow + LTPL_VEL_LANES - 1) / LTPL_VEL_LANES * LTPL_VEL_LANES;
    int q;
    if (t < n_follow)
        q = bf.queue[t];
    else if (t >= n_follow_pad && t - n_follow_pad < n_other)
        q = bf.queue[nq + (t - n_follow_pad)];
    else
        return;
    const int b = q % B;
    int st = bf.status[q];
    const int action = bf.action_id[q];
    const int n = bf.path_len[q];
    const size_t pplane = (size_t)nq * dm.p_max;
    const double* __restrict__ px = bf.path + (size_t)q * dm.p_max;
    const double* __restrict__ py = px + pplane;
    const double* __restrict__ kap = px + 3 * pplane;
    const double* __restrict__ el = px + 4 * pplane;
    double* sc0 = bf.vel_scratch + (size_t)q * dm.p_max;
    double* sc1 = sc0 + pplane;
    double* sc2 = sc1 + pplane;
    double* s = bf.s_vx_ax + (size_t)q * dm.p_max;
    double* vx = s + pplane;   // holds w = v^2 until the final conversion
    double* ax = vx + pplane;

    const double vel_plan = bf.vel[b];  // __v_start (OTH:595)
    const double vel_est = bf.vel_est[b];
    if (vel_plan > prm.vel_max + 0.1) {  // VPFB:106: brake prefix -> the reference raises further down (see header)
        atomicOr(&bf.sc_flags[b], LTPL_SC_BRAKE_PREFIX);
        return;
    }
    const VelCfg c = make_velcfg(prm, s_axm);

    // s = [0, cumsum(el[:-1])]  (OTH:743)
    LTPL_PH_INIT
    {
        double acc = 0.0;
        s[0] = 0.0;
        for (int i = 1; i < n; ++i) {
            acc += el[i - 1];
            s[i] = acc;
        }
    }
    LTPL_PH(0)
    const bool red = (st & LTPL_ST_REDUCED_HORIZON) != 0;
    bool vel_bound = true;
    double* result = vx;

    if (action == LTPL_ACT_FOLLOW) {  // OTH:763-830
        const double ox = bf.cobj[4 * b], oy = bf.cobj[4 * b + 1], ov = bf.cobj[4 * b + 2];
        const double s_obj = s_coord_open_path(px, py, s, el, n, ox, oy);
        const double s_start = s_coord_open_path(px, py, s, el, n, bf.pos[2 * b], bf.pos[2 * b + 1]);
        const double obj_dist = s_obj - s_start;
        LTPL_PH(1)
        const int fl = follow_profile(lt, prm, c, kap, el, s, n, vel_plan, vel_est, ov, obj_dist, bf.cobj_start[b], sc0,
                                      sc1, sc2, vx);
        if (fl & 1) st |= LTPL_ST_TOO_CLOSE;
        vel_bound = !(fl & 2);
        result = sc0;  // a second profile (reduced horizon) goes to scratch
        LTPL_PH(10)
    }
    if (action != LTPL_ACT_FOLLOW || red) {  // OTH:834-923
        const int nn = bf.n_nodes[q];
        const int* nd = bf.nodes + ((size_t)q * dm.h_max + (nn - 1)) * 2;
        const int end_layer = nd[0], end_node = nd[1];
        int dn = end_node - lt.rl_idx[end_layer];
        if (dn < 0) dn = -dn;
        const double raceline_offset = dn * lt.lat_offset;  // quirk q3
        double v_end;
        int v_idx;
        if (red) {
            v_end = 0.0;
            double spl_len = 0.0;
            for (int i = 0; i < n - 1; ++i) spl_len += el[i];
            int first = 0;
            double acc = 0.0;
            for (int i = 0; i < n - 1; ++i) {
                acc += el[i];
                if (!(acc < (spl_len - 5.0))) {
                    first = i;
                    break;
                }
            }
            v_idx = first + 1;
            if (v_idx == 1 && n > 1) v_idx = n;
        } else {
            v_end = lt.vel_rl[end_layer];
            v_end -= fmin(v_end * lt.vel_decrease_lat * raceline_offset, v_end);
            v_idx = n;
        }
        double v_first = 0.0;
        if (v_idx > 1) {
            v_first = sqrt(fb_profile_w(kap, el, v_idx, prm.vel_max, vel_plan, true, v_end, c, result));
        } else {
            result[0] = 0.0;
            v_idx = 1;
        }
        for (int i = v_idx; i < n; ++i) result[i] = 0.0;
        vel_bound = fabs(v_first - vel_plan) < prm.v_max_offset;
        if (action == LTPL_ACT_FOLLOW) {
            // quirk q1 (OTH:923): row 5 decides column-wise; only the vx column differs between the two candidates
            if (n >= 6) {
                if (!(vx[5] < result[5]))
                    for (int i = 0; i < n; ++i) vx[i] = result[i];
            }
        }
    }
    LTPL_PH(8)
    // tph.conv_filt(window = 1) is the identity; vx = sqrt(w); ax profile + standstill fix-up (OTH:926-941)
    {
        double w0 = vx[0];
        for (int i = 0; i < n - 1; ++i) {
            const double w1 = vx[i + 1];
            const double a = (w1 - w0) / (2 * (s[i + 1] - s[i]));
            ax[i] = (w0 <= 1e-16 && fabs(a) <= 1e-8) ? -5.0 : a;
            vx[i] = sqrt(w0);
            w0 = w1;
        }
        vx[n - 1] = sqrt(w0);
        ax[n - 1] = 0.0;
    }

    LTPL_PH(9)
    if (!vel_bound) st |= LTPL_ST_VEL_BOUND_VIOL;
    if (vel_bound || action == LTPL_ACT_FOLLOW || action == LTPL_ACT_STRAIGHT) {  // OTH:945-948 (no backup plan yet)
        st |= LTPL_ST_TRAJ_VALID;
        bf.traj_len[q] = min(n, dm.n_export);
        bf.traj_id[q] = prm.traj_base_id + action;
        const int e = atomicAdd(&bf.queue_cnt[2], 1);  // row in the compact export list
        bf.exp_q[e] = q;
        bf.traj_row[q] = e;
    }
    bf.status[q] = st;
}

// (P, 7) rows s, x, y, psi, kappa, vx, ax of every kept trajectory, cut to nmbr_export_points (OTH:941, LTPL:401-406)
__global__ void __launch_bounds__(LTPL_WARPS_PER_CTA_EXPORT * 32)
k_export(const LtplDims dm, const LtplBuffers bf) {
    const int B = dm.batch;
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * LTPL_WARPS_PER_CTA_EXPORT + (threadIdx.x >> 5);  // row of the compact export list
    if (e >= bf.queue_cnt[2]) return;
    const int q = bf.exp_q[e];
    const int n = bf.traj_len[q];
    const size_t pplane = (size_t)LTPL_NSLOT * B * dm.p_max;
    const int cut = bf.trim ? bf.trim[4 * q + 2] : 0;   // stateful tick: the trajectory starts at the cut index (OTH:700)
    const double* pp = bf.path + (size_t)q * dm.p_max + cut;
    const double* sv = bf.s_vx_ax + (size_t)q * dm.p_max;
    float* out = bf.traj + (size_t)e * dm.n_export * 7;
    for (int i = lane; i < n * 7; i += 32) {
        const int r = i / 7, col = i - 7 * r;
        double v;
        if (col == 0)
            v = sv[r];
        else if (col <= 4)
            v = pp[(size_t)(col - 1) * pplane + r];
        else
            v = sv[(size_t)(col - 4) * pplane + r];
        out[i] = (float)v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stand-alone solver over dense arrays (BASELINE.json config 5: 100 k paths x 500 points)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LTPL_VEL_BLOCK)
k_velprofile_dense(const LtplParams prm, const LtplVelBatch vb) {
    __shared__ double s_axm[3 * LTPL_MAX_AXM];
    stage_axm(prm, s_axm);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= vb.n_paths) return;
    const VelCfg c = make_velcfg(prm, s_axm);
    const int n = vb.n_points;
    const double* __restrict__ kap = vb.kappa + (size_t)i * n;
    const double* __restrict__ el = vb.el + (size_t)i * n;
    double* v = vb.vx + (size_t)i * n;
    double* a = vb.ax + (size_t)i * n;
    fb_profile_w(kap, el, n, prm.vel_max, vb.v_start[i], true, vb.v_end[i], c, v);
    double w0 = v[0];
    for (int k = 0; k < n - 1; ++k) {
        const double w1 = v[k + 1];
        a[k] = (w1 - w0) / (2 * el[k]);
        v[k] = sqrt(w0);
        w0 = w1;
    }
    v[n - 1] = sqrt(w0);
    a[n - 1] = 0.0;
}
