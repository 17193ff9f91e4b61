// ltpl_vel.cuh -- forward/backward ggv-limited velocity profiles.
//   device functions: ax_poss (tph.calc_ax_poss), brake_profile (tph.calc_vel_profile_brake), fb_profile
//   (tph.calc_vel_profile, closed=False, loc_gg mode), follow_profile (helper_funcs/calc_vel_profile_follow.py:78-313)
//   kernels: k_vel (OTH.calc_vel_profile per action, OTH:688-1025), k_export (OTH:941 + LTPL:401-406),
//            k_velprofile_dense (stand-alone solver over dense arrays, BASELINE config 5)
// One THREAD per path: the solver is a serial recurrence over the points of one path (SURVEY hard part 4).
#pragma once
#include "ltpl_common.cuh"

struct VelCfg {
    double ax_max, ay_max;  // local gg * gg_scale (VPFB:213-214)
    double exp_, drag, mass;
    const double* axm_v;
    const double* axm_a;
    int n_axm;
};

enum { MODE_ACCEL_FORW = 0, MODE_DECEL_FORW = 1, MODE_DECEL_BACKW = 2 };

__device__ __forceinline__ double radius_of(double kappa) {
    // radii = abs(1 / kappa), inf where kappa == 0
    return (kappa != 0.0) ? fabs(1.0 / kappa) : LTPL_INF;
}

// tph.calc_vel_profile.calc_ax_poss with a single-row ggv (loc_gg mode: no velocity dependence), mu = 1
__device__ __forceinline__ double ax_poss(double v, double radius, double ax_max_tires, double ay_max_tires, int mode,
                                          const VelCfg& c) {
    const double v2 = v * v;
    const double ay_used = v2 / radius;
    const double ratio = ay_used / ay_max_tires;
    const double radicand = 1.0 - ((c.exp_ == 1.0) ? ratio : pow(ratio, c.exp_));
    double avail = 0.0;
    if (radicand > 0.0) avail = ax_max_tires * ((c.exp_ == 1.0) ? radicand : pow(radicand, 1.0 / c.exp_));
    if (mode == MODE_ACCEL_FORW) {
        const double axm = interp_table(v, c.axm_v, c.axm_a, c.n_axm);
        avail = fmin(avail, axm);
    }
    const double ax_drag = -v2 * c.drag / c.mass;
    return (mode == MODE_DECEL_BACKW) ? (avail - ax_drag) : (avail + ax_drag);
}

// tph.calc_vel_profile_brake: v[0] = v_start, forward integration with full braking, zeros after standstill
__device__ __forceinline__ void brake_profile(const double* kap, const double* el, int n, double v_start,
                                              double ax_max, double ay_max, const VelCfg& c, double* v) {
    if (v_start < 0.0) v_start = 0.0;
    double cur = v_start;
    v[0] = cur;
    int i = 0;
    for (; i < n - 1; ++i) {
        const double a = ax_poss(cur, radius_of(kap[i]), -ax_max, ay_max, MODE_DECEL_FORW, c);
        const double rad = cur * cur + 2 * a * el[i];
        if (rad < 0.0) break;
        cur = sqrt(rad);
        v[i + 1] = cur;
    }
    for (int k = i + 1; k < n; ++k) v[k] = 0.0;
}

// tph.calc_vel_profile(closed=False): initial profile sqrt(ay_max * radius) clipped to v_max, forward acceleration
// phases started at the rising edges of the INITIAL profile, v_end clamp, backward deceleration phases with one
// look-ahead correction (tph __solver_fb_unclosed / __solver_fb_acc_profile).  Single forward + single backward scan:
// "is index i the start of an acceleration phase" only needs the original values at i-1, i, i+1, which are still
// unmodified when the scan reaches i.
// Returns the final v[0] (callers test it against the planned start velocity, OTH:907).
__device__ __forceinline__ double fb_profile(const double* kap, const double* el, int n, double v_max, double v_start,
                                             bool has_end, double v_end, const VelCfg& c, double* v) {
    if (v_start < 0.0) v_start = 0.0;
    if (has_end && v_end < 0.0) v_end = 0.0;
    // ---- forward ----
    double o_i = sqrt(c.ay_max * radius_of(kap[0]));
    if (o_i > v_max) o_i = v_max;
    if (o_i > v_start) o_i = v_start;
    double cur = o_i;
    v[0] = cur;
    bool prev_rise = false, active = false;
    for (int i = 0; i < n - 1; ++i) {
        double o_n = sqrt(c.ay_max * radius_of(kap[i + 1]));
        if (o_n > v_max) o_n = v_max;
        const bool rise = (o_n - o_i) > 0.0;
        if (!active && rise && !prev_rise) active = true;
        double nxt = o_n;
        if (active) {
            const double a = ax_poss(cur, radius_of(kap[i]), c.ax_max, c.ay_max, MODE_ACCEL_FORW, c);
            const double vn = sqrt(cur * cur + 2 * a * el[i]);
            if (vn < o_n) nxt = vn;
            if (vn > v_max) active = false;
        }
        v[i + 1] = nxt;
        cur = nxt;
        prev_rise = rise;
        o_i = o_n;
    }
    if (has_end && v[n - 1] > v_end) v[n - 1] = v_end;
    // ---- backward (flipped arrays, mode decel_backw) ----
    o_i = v[n - 1];
    cur = o_i;
    prev_rise = false;
    active = false;
    for (int j = 0; j < n - 1; ++j) {
        const int p = n - 1 - j, pn = p - 1;
        const double o_n = v[pn];
        const bool rise = (o_n - o_i) > 0.0;
        if (!active && rise && !prev_rise) active = true;
        double nxt = o_n;
        if (active) {
            const double a = ax_poss(cur, radius_of(kap[p]), c.ax_max, c.ay_max, MODE_DECEL_BACKW, c);
            double vn = sqrt(cur * cur + 2 * a * el[pn]);
            const double a2 = ax_poss(vn, radius_of(kap[pn]), c.ax_max, c.ay_max, MODE_DECEL_BACKW, c);
            const double vt = sqrt(cur * cur + 2 * a2 * el[pn]);
            if (vt < vn) vn = vt;
            if (vn < o_n) nxt = vn;
            if (vn > v_max) active = false;
        }
        v[pn] = nxt;
        cur = nxt;
        prev_rise = rise;
        o_i = o_n;
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

// calc_vel_profile_follow (CVPF:78-313).  kap / el / s have n entries (el[n-1] == 0).  vb, prof, compl: n-entry scratch.
// returns flags: bit0 too_close, bit1 vel_bound violated; result (np.minimum(prof, compl)) is written to `out`.
__device__ __forceinline__ int follow_profile(const LatDev& lt, const LtplParams& prm, const VelCfg& c,
                                              const double* kap, const double* el, const double* s, int n,
                                              double v_start, double v_ego, double v_obj, double obj_dist,
                                              double obj_x, double obj_y, double* vb, double* prof, double* compl_,
                                              double* out) {
    int flags = 0;
    const double v_max = prm.vel_max;
    const double control_d = prm.follow_c_p * prm.safety_d + lt.veh_length;
    const double safety_d = prm.safety_d + lt.veh_length;
    if ((obj_dist - safety_d) < 0) flags |= 1;

    // ego brake profile on the local path (CVPF:152-165)
    brake_profile(kap, el, n, v_start, c.ax_max, c.ay_max, c, vb);
    int id_brake = 0;
    while (id_brake < n && vb[id_brake] > 0.1) ++id_brake;
    double ego_stop_dist = 0.0;
    for (int i = 0; i < id_brake; ++i) ego_stop_dist += el[i];

    // opponent matched to the (closed) global race line, rolled to start at its position (CVPF:166-179)
    const int ng = lt.n_glob - 1;
    const double* G = lt.glob_rl;
    int start;
    {
        double bv = LTPL_INF;
        int nb = 0;
        for (int i = 0; i < ng; ++i) {
            const double d = dist2_rn(G[6 * i + 1], G[6 * i + 2], obj_x, obj_y);
            if (d < bv) {
                bv = d;
                nb = i;
            }
        }
        const int idx1 = (nb - 1 < 0) ? ng - 1 : nb - 1;
        const int idx2 = (nb + 1 > ng - 1) ? 0 : nb + 1;
        const double ang1 = fabs(angle3pt(G[6 * nb + 1], G[6 * nb + 2], obj_x, obj_y, G[6 * idx1 + 1], G[6 * idx1 + 2]));
        const double ang2 = fabs(angle3pt(G[6 * nb + 1], G[6 * nb + 2], obj_x, obj_y, G[6 * idx2 + 1], G[6 * idx2 + 2]));
        start = (ang1 >= ang2) ? idx1 : nb;  // closest_indexes[0]
    }
    // opponent brake profile with ggv = [100, 14, 14] (CVPF:134, 185-199): only the stop distance is needed
    double opp_stop_dist = 0.0;
    {
        double v = fmin(v_obj, G[6 * start + 4]);
        if (v < 0.0) v = 0.0;
        int id = 0;
        bool stopped = false;
        while (id < ng && v > 0.1) {
            int r = start + id;
            if (r >= ng) r -= ng;
            opp_stop_dist += G[6 * r + 5];
            ++id;
            if (id <= ng - 1 && !stopped) {
                const double a = ax_poss(v, radius_of(G[6 * r + 3]), -14.0, 14.0, MODE_DECEL_FORW, c);
                const double rad = v * v + 2 * a * G[6 * r + 5];
                if (rad < 0.0) {
                    stopped = true;
                    v = 0.0;
                } else {
                    v = sqrt(rad);
                }
            } else {
                v = 0.0;
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

    const double* src = vb;
    if (ego_stop_dist < s_stop) {
        int idx_c;
        double vcs;
        if (v_start > v_control && stop_idx >= 2) {
            int first = 0;
            for (int i = 0; i < n; ++i)
                if (vb[i] <= v_control) {
                    first = i;
                    break;
                }
            idx_c = min(first, stop_idx);
            if (idx_c == 0) idx_c = stop_idx;
            vcs = vb[idx_c];
        } else {
            if (!(stop_idx >= 2)) flags |= 2;
            idx_c = 0;
            vcs = v_start;
        }
        for (int i = 0; i < idx_c; ++i) prof[i] = vb[i];
        double v0c = vcs;
        if (stop_idx - idx_c > 0) {
            v0c = fb_profile(kap + idx_c, el + idx_c, stop_idx - idx_c + 1, v_control, vcs, true, v_end, c,
                             prof + idx_c);
            if (fabs(v0c - vcs) > 1.0) flags |= 2;
        } else {
            prof[idx_c] = vcs;
        }
        for (int i = stop_idx + 1; i < n; ++i) prof[i] = 0.0;
        const double prof0 = (idx_c == 0) ? v0c : fmax(v_start, 0.0);
        if (fabs(prof0 - v_start) > 1.0) flags |= 2;
        src = prof;
    }
    // complete (unconstrained) profile and intersection (CVPF:296-310)
    fb_profile(kap, el, n, v_max, v_start, false, 0.0, c, compl_);
    for (int i = 0; i < n; ++i) out[i] = fmin(src[i], compl_[i]);
    return flags;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_vel: OTH.get_ref_idx (never planned before, OTH:590-599) + OTH.calc_vel_profile per action (OTH:688-1025)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_vel(const LatDev lt, const LtplParams prm, const LtplDims dm, const LtplBuffers bf) {
    const int B = dm.batch;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= LTPL_NSLOT * B) return;
    const int b = q % B;
    int st = bf.status[q];
    bf.traj_len[q] = 0;
    bf.traj_id[q] = -1;
    if (!(st & LTPL_ST_FOUND)) return;
    const int action = bf.action_id[q];
    const int n = bf.path_len[q];
    const size_t pplane = (size_t)LTPL_NSLOT * B * dm.p_max;
    const double* px = bf.path + (size_t)q * dm.p_max;
    const double* py = px + pplane;
    const double* kap = px + 3 * pplane;
    const double* el = px + 4 * pplane;
    double* sc0 = bf.vel_scratch + (size_t)q * dm.p_max;
    double* sc1 = sc0 + pplane;
    double* sc2 = sc1 + pplane;
    double* s = bf.s_vx_ax + (size_t)q * dm.p_max;
    double* vx = s + pplane;
    double* ax = vx + pplane;

    const double vel_plan = bf.vel[b];  // __v_start (OTH:595)
    const double vel_est = bf.vel_est[b];
    if (vel_plan > prm.vel_max + 0.1) {  // VPFB:106: brake prefix -> the reference raises further down (see header)
        atomicOr(&bf.sc_flags[b], LTPL_SC_BRAKE_PREFIX);
        return;
    }
    VelCfg c;
    c.ax_max = prm.gg_ax * prm.gg_scale;
    c.ay_max = prm.gg_ay * prm.gg_scale;
    c.exp_ = prm.dyn_model_exp;
    c.drag = prm.drag_coeff;
    c.mass = prm.m_veh;
    c.axm_v = prm.axm_v;
    c.axm_a = prm.axm_a;
    c.n_axm = prm.n_axm;

    // s = [0, cumsum(el[:-1])]  (OTH:743)
    {
        double acc = 0.0;
        s[0] = 0.0;
        for (int i = 1; i < n; ++i) {
            acc += el[i - 1];
            s[i] = acc;
        }
    }
    const bool red = (st & LTPL_ST_REDUCED_HORIZON) != 0;
    bool vel_bound = true;
    double* result = vx;

    if (action == LTPL_ACT_FOLLOW) {  // OTH:763-830
        const double ox = bf.cobj[4 * b], oy = bf.cobj[4 * b + 1], ov = bf.cobj[4 * b + 2];
        const double s_obj = s_coord_open_path(px, py, s, el, n, ox, oy);
        const double s_start = s_coord_open_path(px, py, s, el, n, bf.pos[2 * b], bf.pos[2 * b + 1]);
        const double obj_dist = s_obj - s_start;
        const int fl = follow_profile(lt, prm, c, kap, el, s, n, vel_plan, vel_est, ov, obj_dist, ox, oy, sc0, sc1, sc2,
                                      vx);
        if (fl & 1) st |= LTPL_ST_TOO_CLOSE;
        vel_bound = !(fl & 2);
        result = sc0;  // a second profile (reduced horizon) goes to scratch
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
            v_first = fb_profile(kap, el, v_idx, prm.vel_max, vel_plan, true, v_end, c, result);
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
    // tph.conv_filt(window = 1) is the identity; ax profile + standstill fix-up (OTH:926-941)
    for (int i = 0; i < n - 1; ++i) {
        const double a = (vx[i + 1] * vx[i + 1] - vx[i] * vx[i]) / (2 * (s[i + 1] - s[i]));
        ax[i] = (fabs(vx[i]) <= 1e-8 && fabs(a) <= 1e-8) ? -5.0 : a;
    }
    ax[n - 1] = 0.0;

    if (!vel_bound) st |= LTPL_ST_VEL_BOUND_VIOL;
    if (vel_bound || action == LTPL_ACT_FOLLOW || action == LTPL_ACT_STRAIGHT) {  // OTH:945-948 (no backup plan yet)
        st |= LTPL_ST_TRAJ_VALID;
        bf.traj_len[q] = min(n, dm.n_export);
        bf.traj_id[q] = prm.traj_base_id + action;
    }
    bf.status[q] = st;
}

// (P, 7) rows s, x, y, psi, kappa, vx, ax of every kept trajectory, cut to nmbr_export_points (OTH:941, LTPL:401-406)
__global__ void __launch_bounds__(LTPL_WARPS_PER_CTA_EXPORT * 32)
k_export(const LtplDims dm, const LtplBuffers bf) {
    const int B = dm.batch;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * LTPL_WARPS_PER_CTA_EXPORT + (threadIdx.x >> 5);
    if (q >= LTPL_NSLOT * B) return;
    const int n = bf.traj_len[q];
    if (n <= 0) return;
    const size_t pplane = (size_t)LTPL_NSLOT * B * dm.p_max;
    const double* pp = bf.path + (size_t)q * dm.p_max;
    const double* sv = bf.s_vx_ax + (size_t)q * dm.p_max;
    float* out = bf.traj + (size_t)q * dm.n_export * 7;
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
__global__ void __launch_bounds__(128)
k_velprofile_dense(const LtplParams prm, const LtplVelBatch vb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= vb.n_paths) return;
    VelCfg c;
    c.ax_max = prm.gg_ax * prm.gg_scale;
    c.ay_max = prm.gg_ay * prm.gg_scale;
    c.exp_ = prm.dyn_model_exp;
    c.drag = prm.drag_coeff;
    c.mass = prm.m_veh;
    c.axm_v = prm.axm_v;
    c.axm_a = prm.axm_a;
    c.n_axm = prm.n_axm;
    const int n = vb.n_points;
    const double* kap = vb.kappa + (size_t)i * n;
    const double* el = vb.el + (size_t)i * n;
    double* v = vb.vx + (size_t)i * n;
    double* a = vb.ax + (size_t)i * n;
    fb_profile(kap, el, n, prm.vel_max, vb.v_start[i], true, vb.v_end[i], c, v);
    for (int k = 0; k < n - 1; ++k) a[k] = (v[k + 1] * v[k + 1] - v[k] * v[k]) / (2 * el[k]);
    a[n - 1] = 0.0;
}
