// ltpl_velprofile.cuh -- k_velprofile_tiled: stand-alone forward/backward ggv solver over dense [n_paths][n_points]
// arrays (BASELINE.json config 5; tph.calc_vel_profile(closed=False, loc_gg mode), oracle/tph_port.py:463-538).
#pragma once
#include "ltpl_vel.cuh"

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

// mode 'decel_backw': tyre - drag
__device__ __forceinline__ double acc_backw(double w, double kabs, const VelCfg& c) {
    return fma(w, c.dm, acc_tire(w, kabs, c.ax_max, c.inv_ay, c.exp_));
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
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
