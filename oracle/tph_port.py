"""
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU restatement (NumPy float64 / pure Python loops) of the functions of the un-vendored third-party dependency
``trajectory_planning_helpers==0.75`` (pinned at /root/reference/requirements.txt:5) that the reference's online
planning path calls.  The package source is NOT part of /root/reference and there is no network, so the published
algorithm is restated here from the upstream sources (from memory); parity of these functions against a real tph
install is therefore *unpinned* (see DESIGN.md "Oracle pinning").  What IS pinned: the reference's own Python files
are executed verbatim on top of this module (oracle/shims + oracle/gen_golden.py) to produce tests/golden/*.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

Reference call sites (file:line relative to /root/reference):
  calc_splines            MOPG:305-309, OTH:244-246, offline_graph/src/gen_edges.py:47,88-92
  interp_splines          MOPG:312-316, OTH:248-252, gen_edges.py:128-131
  calc_head_curv_an       MOPG:318-322, OTH:254-257, gen_edges.py:133-136
  calc_head_curv_num      offline_graph/src/gen_node_skeleton.py:63-65,86-92
  normalize_psi           gen_node_skeleton.py:142-150
  calc_vel_profile        online_graph/src/VpForwardBackward.py:213-225, helper_funcs/src/calc_vel_profile_follow.py:268,297
  calc_vel_profile_brake  VpForwardBackward.py:115-122,247-253, calc_vel_profile_follow.py:152,185, calc_brake_emergency.py:30
  conv_filt               OTH:928-930,988-990
  calc_ax_profile         OTH:935-936,996-998, calc_brake_emergency.py:39
"""

import math

import numpy as np


# ----------------------------------------------------------------------------------------------------------------------
# normalize_psi
# ----------------------------------------------------------------------------------------------------------------------
def normalize_psi(psi):
    """tph.normalize_psi: map angle(s) to [-pi, pi)."""
    psi_out = np.sign(psi) * np.mod(np.abs(psi), 2 * math.pi)

    if type(psi_out) is np.ndarray:
        psi_out[psi_out >= math.pi] -= 2 * math.pi
        psi_out[psi_out < -math.pi] += 2 * math.pi
    else:
        if psi_out >= math.pi:
            psi_out -= 2 * math.pi
        elif psi_out < -math.pi:
            psi_out += 2 * math.pi

    return psi_out


# ----------------------------------------------------------------------------------------------------------------------
# calc_splines
# ----------------------------------------------------------------------------------------------------------------------
def calc_splines(path, el_lengths=None, psi_s=None, psi_e=None, use_dist_scaling=True):
    """tph.calc_splines: C2 cubic splines x(t), y(t), t in [0, 1] per segment, dense (4N x 4N) LES."""
    # closed iff first == last point and no start heading given
    if np.all(np.isclose(path[0], path[-1])) and psi_s is None:
        closed = True
    else:
        closed = False

    if not closed and (psi_s is None or psi_e is None):
        raise RuntimeError("Headings must be provided for unclosed spline calculation!")

    if el_lengths is not None and path.shape[0] != el_lengths.size + 1:
        raise RuntimeError("el_lengths input must be one element smaller than path input!")

    if use_dist_scaling and el_lengths is None:
        el_lengths = np.sqrt(np.sum(np.power(np.diff(path, axis=0), 2), axis=1))
    elif el_lengths is not None:
        el_lengths = np.copy(el_lengths)

    if use_dist_scaling and closed:
        el_lengths = np.append(el_lengths, el_lengths[0])

    no_splines = path.shape[0] - 1

    if use_dist_scaling:
        scaling = el_lengths[:-1] / el_lengths[1:]
    else:
        scaling = np.ones(no_splines - 1)

    M = np.zeros((no_splines * 4, no_splines * 4))
    b_x = np.zeros((no_splines * 4, 1))
    b_y = np.zeros((no_splines * 4, 1))

    template_M = np.array([[1, 0, 0, 0, 0, 0, 0, 0],
                           [1, 1, 1, 1, 0, 0, 0, 0],
                           [0, 1, 2, 3, 0, -1, 0, 0],
                           [0, 0, 2, 6, 0, 0, -2, 0]])

    for i in range(no_splines):
        j = i * 4

        if i < no_splines - 1:
            M[j: j + 4, j: j + 8] = template_M
            M[j + 2, j + 5] *= scaling[i]
            M[j + 3, j + 6] *= math.pow(scaling[i], 2)
        else:
            M[j: j + 2, j: j + 4] = [[1, 0, 0, 0],
                                     [1, 1, 1, 1]]

        b_x[j: j + 2] = [[path[i, 0]],
                         [path[i + 1, 0]]]
        b_y[j: j + 2] = [[path[i, 1]],
                         [path[i + 1, 1]]]

    if not closed:
        # heading start point (t = 0 of first spline)
        M[-2, 1] = 1

        if el_lengths is None:
            el_length_s = 1.0
        else:
            el_length_s = el_lengths[0]

        b_x[-2] = math.cos(psi_s + math.pi / 2) * el_length_s
        b_y[-2] = math.sin(psi_s + math.pi / 2) * el_length_s

        # heading end point (t = 1 of last spline)
        M[-1, -4:] = [0, 1, 2, 3]

        if el_lengths is None:
            el_length_e = 1.0
        else:
            el_length_e = el_lengths[-1]

        b_x[-1] = math.cos(psi_e + math.pi / 2) * el_length_e
        b_y[-1] = math.sin(psi_e + math.pi / 2) * el_length_e

    else:
        # heading continuity last -> first spline
        M[-2, 1] = scaling[-1]
        M[-2, -3:] = [-1, -2, -3]

        # curvature continuity last -> first spline
        M[-1, 2] = 2 * math.pow(scaling[-1], 2)
        M[-1, -2:] = [-2, -6]

    x_les = np.squeeze(np.linalg.solve(M, b_x))
    y_les = np.squeeze(np.linalg.solve(M, b_y))

    coeffs_x = np.reshape(x_les, (no_splines, 4))
    coeffs_y = np.reshape(y_les, (no_splines, 4))

    normvec = np.stack((coeffs_y[:, 1], -coeffs_x[:, 1]), axis=1)
    norm_factors = 1.0 / np.sqrt(np.sum(np.power(normvec, 2), axis=1))
    normvec_normalized = np.expand_dims(norm_factors, axis=1) * normvec

    return coeffs_x, coeffs_y, M, normvec_normalized


# ----------------------------------------------------------------------------------------------------------------------
# calc_spline_lengths
# ----------------------------------------------------------------------------------------------------------------------
def calc_spline_lengths(coeffs_x, coeffs_y, quickndirty=False, no_interp_points=15):
    """tph.calc_spline_lengths: polyline length over `no_interp_points` equidistant t samples per spline."""
    if coeffs_x.size == 4 and coeffs_x.shape[0] == 4:
        coeffs_x = np.expand_dims(coeffs_x, 0)
        coeffs_y = np.expand_dims(coeffs_y, 0)

    no_splines = coeffs_x.shape[0]
    spline_lengths = np.zeros(no_splines)

    if quickndirty:
        for i in range(no_splines):
            spline_lengths[i] = math.sqrt(math.pow(np.sum(coeffs_x[i]) - coeffs_x[i, 0], 2)
                                          + math.pow(np.sum(coeffs_y[i]) - coeffs_y[i, 0], 2))
    else:
        t_steps = np.linspace(0.0, 1.0, no_interp_points)
        spl_coords = np.zeros((no_interp_points, 2))

        for i in range(no_splines):
            spl_coords[:, 0] = coeffs_x[i, 0] \
                + coeffs_x[i, 1] * t_steps \
                + coeffs_x[i, 2] * np.power(t_steps, 2) \
                + coeffs_x[i, 3] * np.power(t_steps, 3)
            spl_coords[:, 1] = coeffs_y[i, 0] \
                + coeffs_y[i, 1] * t_steps \
                + coeffs_y[i, 2] * np.power(t_steps, 2) \
                + coeffs_y[i, 3] * np.power(t_steps, 3)

            spline_lengths[i] = np.sum(np.sqrt(np.sum(np.power(np.diff(spl_coords, axis=0), 2), axis=1)))

    return spline_lengths


# ----------------------------------------------------------------------------------------------------------------------
# interp_splines
# ----------------------------------------------------------------------------------------------------------------------
def interp_splines(coeffs_x, coeffs_y, spline_lengths=None, incl_last_point=False, stepsize_approx=None,
                   stepnum_fixed=None):
    """tph.interp_splines: sample splines either ~equidistantly (stepsize_approx) or with fixed counts per spline."""
    if coeffs_x.shape[0] != coeffs_y.shape[0]:
        raise RuntimeError("Coefficient matrices must have the same length!")

    if spline_lengths is not None and coeffs_x.shape[0] != spline_lengths.size:
        raise RuntimeError("coeffs_x/y and spline_lengths must have the same length!")

    if not (coeffs_x.ndim == 2 and coeffs_y.ndim == 2):
        raise RuntimeError("Coefficient matrices do not have two dimensions!")

    if (stepsize_approx is None and stepnum_fixed is None) \
            or (stepsize_approx is not None and stepnum_fixed is not None):
        raise RuntimeError("Provide one of 'stepsize_approx' and 'stepnum_fixed' and set the other to 'None'!")

    if stepnum_fixed is not None and len(stepnum_fixed) != coeffs_x.shape[0]:
        raise RuntimeError("The provided list 'stepnum_fixed' must hold an entry for every spline!")

    if stepsize_approx is not None:
        if spline_lengths is None:
            spline_lengths = calc_spline_lengths(coeffs_x=coeffs_x, coeffs_y=coeffs_y, quickndirty=False)

        dists_cum = np.cumsum(spline_lengths)

        no_interp_points = math.ceil(dists_cum[-1] / stepsize_approx) + 1
        dists_interp = np.linspace(0.0, dists_cum[-1], no_interp_points)
    else:
        no_interp_points = sum(stepnum_fixed) - (len(stepnum_fixed) - 1)
        dists_interp = None

    path_interp = np.zeros((no_interp_points, 2))
    spline_inds = np.zeros(no_interp_points, dtype=int)
    t_values = np.zeros(no_interp_points)

    if stepsize_approx is not None:
        for i in range(no_interp_points - 1):
            j = np.argmax(dists_interp[i] < dists_cum)
            spline_inds[i] = j

            if j > 0:
                t_values[i] = (dists_interp[i] - dists_cum[j - 1]) / spline_lengths[j]
            else:
                if spline_lengths.ndim == 0:
                    t_values[i] = dists_interp[i] / spline_lengths
                else:
                    t_values[i] = dists_interp[i] / spline_lengths[0]

            path_interp[i, 0] = coeffs_x[j, 0] \
                + coeffs_x[j, 1] * t_values[i] \
                + coeffs_x[j, 2] * math.pow(t_values[i], 2) \
                + coeffs_x[j, 3] * math.pow(t_values[i], 3)

            path_interp[i, 1] = coeffs_y[j, 0] \
                + coeffs_y[j, 1] * t_values[i] \
                + coeffs_y[j, 2] * math.pow(t_values[i], 2) \
                + coeffs_y[j, 3] * math.pow(t_values[i], 3)
    else:
        j = 0

        for i in range(len(stepnum_fixed)):
            if i < len(stepnum_fixed) - 1:
                t_values[j:(j + stepnum_fixed[i] - 1)] = np.linspace(0, 1, stepnum_fixed[i])[:-1]
                spline_inds[j:(j + stepnum_fixed[i] - 1)] = i
                j += stepnum_fixed[i] - 1
            else:
                t_values[j:(j + stepnum_fixed[i])] = np.linspace(0, 1, stepnum_fixed[i])
                spline_inds[j:(j + stepnum_fixed[i])] = i
                j += stepnum_fixed[i]

        t_set = np.column_stack((np.ones(no_interp_points), t_values, np.power(t_values, 2), np.power(t_values, 3)))

        n_samples = np.array(stepnum_fixed)
        n_samples[:-1] -= 1

        path_interp[:, 0] = np.sum(np.multiply(np.repeat(coeffs_x, n_samples, axis=0), t_set), axis=1)
        path_interp[:, 1] = np.sum(np.multiply(np.repeat(coeffs_y, n_samples, axis=0), t_set), axis=1)

    if incl_last_point:
        path_interp[-1, 0] = np.sum(coeffs_x[-1])
        path_interp[-1, 1] = np.sum(coeffs_y[-1])
        spline_inds[-1] = coeffs_x.shape[0] - 1
        t_values[-1] = 1.0
    else:
        path_interp = path_interp[:-1]
        spline_inds = spline_inds[:-1]
        t_values = t_values[:-1]

        if dists_interp is not None:
            dists_interp = dists_interp[:-1]

    return path_interp, spline_inds, t_values, dists_interp


# ----------------------------------------------------------------------------------------------------------------------
# calc_head_curv_an
# ----------------------------------------------------------------------------------------------------------------------
def calc_head_curv_an(coeffs_x, coeffs_y, ind_spls, t_spls, calc_curv=True, calc_dcurv=False):
    """tph.calc_head_curv_an: analytic heading (0 = north) and curvature on cubic splines."""
    if coeffs_x.shape[0] != coeffs_y.shape[0]:
        raise ValueError("Coefficient matrices must have the same length!")

    if ind_spls.size != t_spls.size:
        raise ValueError("ind_spls and t_spls must have the same length!")

    x_d = coeffs_x[ind_spls, 1] \
        + 2 * coeffs_x[ind_spls, 2] * t_spls \
        + 3 * coeffs_x[ind_spls, 3] * np.power(t_spls, 2)

    y_d = coeffs_y[ind_spls, 1] \
        + 2 * coeffs_y[ind_spls, 2] * t_spls \
        + 3 * coeffs_y[ind_spls, 3] * np.power(t_spls, 2)

    psi = np.arctan2(y_d, x_d) - math.pi / 2
    psi = normalize_psi(psi)

    if calc_curv:
        x_dd = 2 * coeffs_x[ind_spls, 2] \
            + 6 * coeffs_x[ind_spls, 3] * t_spls

        y_dd = 2 * coeffs_y[ind_spls, 2] \
            + 6 * coeffs_y[ind_spls, 3] * t_spls

        kappa = (x_d * y_dd - y_d * x_dd) / np.power(np.power(x_d, 2) + np.power(y_d, 2), 1.5)
    else:
        kappa = 0.0

    if calc_dcurv:
        x_ddd = 6 * coeffs_x[ind_spls, 3]
        y_ddd = 6 * coeffs_y[ind_spls, 3]
        dkappa = ((np.power(x_d, 2) + np.power(y_d, 2)) * (x_d * y_ddd - y_d * x_ddd)
                  - 3 * (x_d * y_dd - y_d * x_dd) * (x_d * x_dd + y_d * y_dd)) \
            / np.power(np.power(x_d, 2) + np.power(y_d, 2), 3)
        return psi, kappa, dkappa

    return psi, kappa


# ----------------------------------------------------------------------------------------------------------------------
# calc_head_curv_num (heading part is what the offline pipeline uses; curvature restated for completeness)
# ----------------------------------------------------------------------------------------------------------------------
def calc_head_curv_num(path, el_lengths, is_closed, stepsize_psi_preview=1.0, stepsize_psi_review=1.0,
                       stepsize_curv_preview=2.0, stepsize_curv_review=2.0, calc_curv=True):
    """tph.calc_head_curv_num: numerical heading/curvature via preview/review secants."""
    if is_closed and path.shape[0] != el_lengths.size:
        raise RuntimeError("path and el_lenghts must have the same length!")
    elif not is_closed and path.shape[0] != el_lengths.size + 1:
        raise RuntimeError("path must have the length of el_lengths + 1!")

    no_points = path.shape[0]

    if is_closed:
        ind_step_preview_psi = round(stepsize_psi_preview / float(np.average(el_lengths)))
        ind_step_review_psi = round(stepsize_psi_review / float(np.average(el_lengths)))
        ind_step_preview_curv = round(stepsize_curv_preview / float(np.average(el_lengths)))
        ind_step_review_curv = round(stepsize_curv_review / float(np.average(el_lengths)))

        ind_step_preview_psi = max(ind_step_preview_psi, 1)
        ind_step_review_psi = max(ind_step_review_psi, 1)
        ind_step_preview_curv = max(ind_step_preview_curv, 1)
        ind_step_review_curv = max(ind_step_review_curv, 1)

        steps_tot_psi = ind_step_preview_psi + ind_step_review_psi
        steps_tot_curv = ind_step_preview_curv + ind_step_review_curv

        path_temp = np.vstack((path[-ind_step_review_psi:], path, path[:ind_step_preview_psi]))
        tangvecs = np.stack((path_temp[steps_tot_psi:, 0] - path_temp[:-steps_tot_psi, 0],
                             path_temp[steps_tot_psi:, 1] - path_temp[:-steps_tot_psi, 1]), axis=1)

        psi = np.arctan2(tangvecs[:, 1], tangvecs[:, 0]) - math.pi / 2
        psi = normalize_psi(psi)

        if calc_curv:
            psi_temp = np.insert(psi, 0, psi[-ind_step_review_curv:])
            psi_temp = np.append(psi_temp, psi[:ind_step_preview_curv])

            delta_psi = np.zeros(no_points)

            for i in range(no_points):
                delta_psi[i] = normalize_psi(psi_temp[i + steps_tot_curv] - psi_temp[i])

            s_points_cl = np.cumsum(el_lengths)
            s_points_cl = np.insert(s_points_cl, 0, 0.0)
            s_points = s_points_cl[:-1]
            s_points_cl_reverse = np.flipud(-np.cumsum(np.flipud(el_lengths)))

            s_points_temp = np.insert(s_points, 0, s_points_cl_reverse[-ind_step_review_curv:])
            s_points_temp = np.append(s_points_temp, s_points_cl[-1] + s_points[:ind_step_preview_curv])

            kappa = delta_psi / (s_points_temp[steps_tot_curv:] - s_points_temp[:-steps_tot_curv])
        else:
            kappa = 0.0
    else:
        # heading (unclosed): central differences inside, one-sided at the ends
        tangvecs = np.zeros((no_points, 2))
        tangvecs[0, 0] = path[1, 0] - path[0, 0]
        tangvecs[0, 1] = path[1, 1] - path[0, 1]
        tangvecs[1:-1, 0] = path[2:, 0] - path[:-2, 0]
        tangvecs[1:-1, 1] = path[2:, 1] - path[:-2, 1]
        tangvecs[-1, 0] = path[-1, 0] - path[-2, 0]
        tangvecs[-1, 1] = path[-1, 1] - path[-2, 1]

        psi = np.arctan2(tangvecs[:, 1], tangvecs[:, 0]) - math.pi / 2
        psi = normalize_psi(psi)

        if calc_curv:
            delta_psi = np.zeros(no_points)
            delta_psi[0] = psi[1] - psi[0]
            delta_psi[1:-1] = psi[2:] - psi[:-2]
            delta_psi[-1] = psi[-1] - psi[-2]
            delta_psi = normalize_psi(delta_psi)

            kappa = np.zeros(no_points)
            kappa[0] = delta_psi[0] / el_lengths[0]
            kappa[1:-1] = delta_psi[1:-1] / (el_lengths[1:] + el_lengths[:-1])
            kappa[-1] = delta_psi[-1] / el_lengths[-1]
        else:
            kappa = 0.0

    return psi, kappa


# ----------------------------------------------------------------------------------------------------------------------
# calc_ax_poss / calc_vel_profile (forward-backward solver)
# ----------------------------------------------------------------------------------------------------------------------
def calc_ax_poss(vx_start, radius, ggv, mu, dyn_model_exp, drag_coeff, m_veh, ax_max_machines=None,
                 mode='accel_forw'):
    """tph.calc_vel_profile.calc_ax_poss: possible longitudinal acceleration at (v, radius)."""
    if mode not in ['accel_forw', 'decel_forw', 'decel_backw']:
        raise RuntimeError("Unknown operation mode for calc_ax_poss!")

    if mode == 'accel_forw' and ax_max_machines is None:
        raise RuntimeError("ax_max_machines is required if operation mode is accel_forw!")

    if ggv.ndim != 2 or ggv.shape[1] != 3:
        raise RuntimeError("ggv must have two dimensions and three columns [vx, ax_max, ay_max]!")

    # tire potential
    ax_max_tires = mu * np.interp(vx_start, ggv[:, 0], ggv[:, 1])
    ay_max_tires = mu * np.interp(vx_start, ggv[:, 0], ggv[:, 2])
    ay_used = math.pow(vx_start, 2) / radius

    if mode in ['accel_forw', 'decel_backw'] and ax_max_tires < 0.0:
        print("WARNING: Inverting sign of ax_max_tires because it should be positive but was negative!")
        ax_max_tires *= -1.0
    elif mode == 'decel_forw' and ax_max_tires > 0.0:
        print("WARNING: Inverting sign of ax_max_tires because it should be negative but was positve!")
        ax_max_tires *= -1.0

    radicand = 1.0 - math.pow(ay_used / ay_max_tires, dyn_model_exp)

    if radicand > 0.0:
        ax_avail_tires = ax_max_tires * math.pow(radicand, 1.0 / dyn_model_exp)
    else:
        ax_avail_tires = 0.0

    # machine limits (forward acceleration only)
    if mode == 'accel_forw':
        ax_max_machines_tmp = np.interp(vx_start, ax_max_machines[:, 0], ax_max_machines[:, 1])
        ax_avail_vehicle = min(ax_avail_tires, ax_max_machines_tmp)
    else:
        ax_avail_vehicle = ax_avail_tires

    # drag
    ax_drag = -math.pow(vx_start, 2) * drag_coeff / m_veh

    if mode in ['accel_forw', 'decel_forw']:
        ax_final = ax_avail_vehicle + ax_drag
    else:
        ax_final = ax_avail_vehicle - ax_drag

    return ax_final


def _solver_fb_acc_profile(p_ggv, ax_max_machines, v_max, radii, el_lengths, mu, vx_profile, dyn_model_exp,
                           drag_coeff, m_veh, backwards=False):
    no_points = vx_profile.size

    if backwards:
        radii_mod = np.flipud(radii)
        el_lengths_mod = np.flipud(el_lengths)
        mu_mod = np.flipud(mu)
        vx_profile = np.flipud(vx_profile)
        mode = 'decel_backw'
    else:
        radii_mod = radii
        el_lengths_mod = el_lengths
        mu_mod = mu
        mode = 'accel_forw'

    # start points of acceleration phases
    vx_diffs = np.diff(vx_profile)
    acc_inds = np.where(vx_diffs > 0.0)[0]
    if acc_inds.size != 0:
        acc_inds_diffs = np.diff(acc_inds)
        acc_inds_diffs = np.insert(acc_inds_diffs, 0, 2)
        acc_inds_rel = acc_inds[acc_inds_diffs > 1]
    else:
        acc_inds_rel = []

    acc_inds_rel = list(acc_inds_rel)

    while acc_inds_rel:
        i = acc_inds_rel.pop(0)

        while i < no_points - 1:
            ax_possible_cur = calc_ax_poss(vx_start=vx_profile[i],
                                           radius=radii_mod[i],
                                           ggv=p_ggv[i],
                                           ax_max_machines=ax_max_machines,
                                           mu=mu_mod[i],
                                           mode=mode,
                                           dyn_model_exp=dyn_model_exp,
                                           drag_coeff=drag_coeff,
                                           m_veh=m_veh)

            vx_possible_next = math.sqrt(math.pow(vx_profile[i], 2) + 2 * ax_possible_cur * el_lengths_mod[i])

            if backwards:
                # one look-ahead correction (ax evaluated at the next point's velocity / radius)
                for j in range(1):
                    ax_possible_next = calc_ax_poss(vx_start=vx_possible_next,
                                                    radius=radii_mod[i + 1],
                                                    ggv=p_ggv[i + 1],
                                                    ax_max_machines=ax_max_machines,
                                                    mu=mu_mod[i + 1],
                                                    mode=mode,
                                                    dyn_model_exp=dyn_model_exp,
                                                    drag_coeff=drag_coeff,
                                                    m_veh=m_veh)

                    vx_tmp = math.sqrt(math.pow(vx_profile[i], 2) + 2 * ax_possible_next * el_lengths_mod[i])

                    if vx_tmp < vx_possible_next:
                        vx_possible_next = vx_tmp
                    else:
                        break

            if vx_possible_next < vx_profile[i + 1]:
                vx_profile[i + 1] = vx_possible_next

            i += 1

            if vx_possible_next > v_max or (acc_inds_rel and i >= acc_inds_rel[0]):
                break

    if backwards:
        vx_profile = np.flipud(vx_profile)

    return vx_profile


def _solver_fb_unclosed(p_ggv, ax_max_machines, v_max, radii, el_lengths, v_start, drag_coeff, m_veh, op_mode,
                        mu=None, v_end=None, dyn_model_exp=1.0):
    if mu is None:
        mu = np.ones(radii.size)
        mu_mean = 1.0
    else:
        mu_mean = np.mean(mu)

    if op_mode == 'ggv':
        ay_max_global = mu_mean * np.amin(p_ggv[0, :, 2])
        vx_profile = np.sqrt(ay_max_global * radii)

        ay_max_curr = mu * np.interp(vx_profile, p_ggv[0, :, 0], p_ggv[0, :, 2])
        vx_profile = np.sqrt(np.multiply(ay_max_curr, radii))
    else:
        vx_profile = np.sqrt(p_ggv[:, 0, 2] * radii)

    vx_profile[vx_profile > v_max] = v_max

    if vx_profile[0] > v_start:
        vx_profile[0] = v_start

    vx_profile = _solver_fb_acc_profile(p_ggv=p_ggv, ax_max_machines=ax_max_machines, v_max=v_max, radii=radii,
                                        el_lengths=el_lengths, mu=mu, vx_profile=vx_profile, backwards=False,
                                        dyn_model_exp=dyn_model_exp, drag_coeff=drag_coeff, m_veh=m_veh)

    if v_end is not None and vx_profile[-1] > v_end:
        vx_profile[-1] = v_end

    vx_profile = _solver_fb_acc_profile(p_ggv=p_ggv, ax_max_machines=ax_max_machines, v_max=v_max, radii=radii,
                                        el_lengths=el_lengths, mu=mu, vx_profile=vx_profile, backwards=True,
                                        dyn_model_exp=dyn_model_exp, drag_coeff=drag_coeff, m_veh=m_veh)

    return vx_profile


def calc_vel_profile(ax_max_machines, kappa, el_lengths, closed, drag_coeff, m_veh, ggv=None, loc_gg=None,
                     v_max=None, dyn_model_exp=1.0, mu=None, v_start=None, v_end=None, filt_window=None):
    """tph.calc_vel_profile (only the unclosed solver is on the reference's online path)."""
    if (ggv is not None or mu is not None) and loc_gg is not None:
        raise RuntimeError("Either ggv and optionally mu OR loc_gg must be supplied, not both (or all) of them!")

    if ggv is None and loc_gg is None:
        raise RuntimeError("Either ggv or loc_gg must be supplied!")

    if loc_gg is not None:
        if loc_gg.ndim != 2:
            raise RuntimeError("loc_gg must have two dimensions!")
        if loc_gg.shape[0] != kappa.size:
            raise RuntimeError("Length of loc_gg and kappa must be equal!")
        if loc_gg.shape[1] != 2:
            raise RuntimeError("loc_gg must consist of two columns: [ax_max, ay_max]!")

    if ggv is not None and ggv.shape[1] != 3:
        raise RuntimeError("ggv diagram must consist of the three columns [vx, ax_max, ay_max]!")

    if mu is not None and kappa.size != mu.size:
        raise RuntimeError("kappa and mu must have the same length!")

    if closed and kappa.size != el_lengths.size:
        raise RuntimeError("kappa and el_lengths must have the same length if closed!")
    elif not closed and kappa.size != el_lengths.size + 1:
        raise RuntimeError("kappa must have the length of el_lengths + 1 if unclosed!")

    if not closed and v_start is None:
        raise RuntimeError("v_start must be provided for the unclosed case!")

    if v_start is not None and v_start < 0.0:
        v_start = 0.0
        print('WARNING: Input v_start was < 0.0. Using v_start = 0.0 instead!')

    if v_end is not None and v_end < 0.0:
        v_end = 0.0
        print('WARNING: Input v_end was < 0.0. Using v_end = 0.0 instead!')

    if not 1.0 <= dyn_model_exp <= 2.0:
        print('WARNING: Exponent for the vehicle dynamics model should be in the range [1.0, 2.0]!')

    if ax_max_machines.shape[1] != 2:
        raise RuntimeError("ax_max_machines must consist of the two columns [vx, ax_max_machines]!")

    if v_max is None:
        if ggv is None:
            raise RuntimeError("v_max must be supplied if ggv is None!")
        else:
            v_max = min(ggv[-1, 0], ax_max_machines[-1, 0])
    else:
        if ggv is not None and ggv[-1, 0] < v_max:
            raise RuntimeError("ggv has to cover the entire velocity range of the car (i.e. >= v_max)!")
        if ax_max_machines[-1, 0] < v_max:
            raise RuntimeError("ax_max_machines has to cover the entire velocity range of the car (i.e. >= v_max)!")

    if ggv is not None:
        p_ggv = np.repeat(np.expand_dims(ggv, axis=0), kappa.size, axis=0)
        op_mode = 'ggv'
    else:
        p_ggv = np.expand_dims(np.column_stack((np.ones(loc_gg.shape[0]) * 10.0, loc_gg)), axis=1)
        op_mode = 'loc_gg'

    radii = np.abs(np.divide(1.0, kappa, out=np.full(kappa.size, np.inf), where=kappa != 0.0))

    if not closed:
        vx_profile = _solver_fb_unclosed(p_ggv=p_ggv, ax_max_machines=ax_max_machines, v_max=v_max, radii=radii,
                                         el_lengths=el_lengths, mu=mu, v_start=v_start, v_end=v_end,
                                         dyn_model_exp=dyn_model_exp, drag_coeff=drag_coeff, m_veh=m_veh,
                                         op_mode=op_mode)
    else:
        raise NotImplementedError("closed fb solver is not on the reference's online path (VPFB:224 closed=False)")

    if filt_window is not None:
        vx_profile = conv_filt(signal=vx_profile, filt_window=filt_window, closed=closed)

    return vx_profile


# ----------------------------------------------------------------------------------------------------------------------
# calc_vel_profile_brake
# ----------------------------------------------------------------------------------------------------------------------
def calc_vel_profile_brake(kappa, el_lengths, v_start, drag_coeff, m_veh, ggv=None, loc_gg=None, dyn_model_exp=1.0,
                           mu=None, decel_max=None):
    """tph.calc_vel_profile_brake: pure forward maximum-braking profile (zeros after standstill)."""
    if (ggv is not None or mu is not None) and loc_gg is not None:
        raise RuntimeError("Either ggv and optionally mu OR loc_gg must be supplied, not both (or all) of them!")

    if ggv is None and loc_gg is None:
        raise RuntimeError("Either ggv or loc_gg must be supplied!")

    if loc_gg is not None:
        if loc_gg.ndim != 2:
            raise RuntimeError("loc_gg must have two dimensions!")
        if loc_gg.shape[0] != kappa.size:
            raise RuntimeError("Length of loc_gg and kappa must be equal!")
        if loc_gg.shape[1] != 2:
            raise RuntimeError("loc_gg must consist of two columns: [ax_max, ay_max]!")

    if ggv is not None and ggv.shape[1] != 3:
        raise RuntimeError("ggv diagram must consist of the three columns [vx, ax_max, ay_max]!")

    if mu is not None and kappa.size != mu.size:
        raise RuntimeError("kappa and mu must have the same length!")

    if kappa.size != el_lengths.size + 1:
        raise RuntimeError("kappa must have the length of el_lengths + 1!")

    if v_start < 0.0:
        v_start = 0.0
        print('WARNING: Input v_start was < 0.0. Using v_start = 0.0 instead!')

    if not 1.0 <= dyn_model_exp <= 2.0:
        print('WARNING: Exponent for the vehicle dynamics model should be in the range [1.0, 2.0]!')

    if mu is None:
        mu = np.ones(kappa.size)

    if ggv is not None:
        p_ggv = np.repeat(np.expand_dims(ggv, axis=0), kappa.size, axis=0)
    else:
        p_ggv = np.expand_dims(np.column_stack((np.ones(loc_gg.shape[0]) * 10.0, loc_gg)), axis=1)

    radii = np.abs(np.divide(1, kappa, out=np.full(kappa.size, np.inf), where=kappa != 0))

    vx_profile = np.zeros(kappa.size)
    vx_profile[0] = v_start

    for i in range(vx_profile.size - 1):
        ggv_mod = np.copy(p_ggv[i])
        ggv_mod[:, 1] *= -1.0
        ax_final = calc_ax_poss(vx_start=vx_profile[i],
                                radius=radii[i],
                                ggv=ggv_mod,
                                mu=mu[i],
                                mode='decel_forw',
                                dyn_model_exp=dyn_model_exp,
                                drag_coeff=drag_coeff,
                                m_veh=m_veh)

        ax_drag = -math.pow(vx_profile[i], 2) * drag_coeff / m_veh

        if decel_max is not None and ax_final < decel_max:
            if ax_drag < decel_max:
                ax_final = ax_drag
            else:
                ax_final = decel_max

        radicand = math.pow(vx_profile[i], 2) + 2 * ax_final * el_lengths[i]

        if radicand < 0.0:
            break
        else:
            vx_profile[i + 1] = math.sqrt(radicand)

    return vx_profile


# ----------------------------------------------------------------------------------------------------------------------
# conv_filt / calc_ax_profile
# ----------------------------------------------------------------------------------------------------------------------
def conv_filt(signal, filt_window, closed):
    """tph.conv_filt: moving average; open signals keep the first/last half-window samples (w=1 => identity)."""
    if not filt_window % 2 == 1:
        raise RuntimeError("Window width of moving average filter must be odd!")

    w_window_half = int((filt_window - 1) / 2)

    if closed:
        signal_tmp = np.concatenate((signal[-w_window_half:], signal, signal[:w_window_half]), axis=0)
        signal_filt = np.convolve(signal_tmp,
                                  np.ones(filt_window) / float(filt_window),
                                  mode="same")[w_window_half:-w_window_half]
    else:
        signal_filt = np.copy(signal)
        signal_filt[w_window_half:-w_window_half] = np.convolve(signal,
                                                                np.ones(filt_window) / float(filt_window),
                                                                mode="same")[w_window_half:-w_window_half]

    return signal_filt


def calc_ax_profile(vx_profile, el_lengths, eq_length_output=False):
    """tph.calc_ax_profile: ax = (v1^2 - v0^2) / (2 ds)."""
    if vx_profile.size != el_lengths.size + 1:
        raise RuntimeError("Array size of vx_profile should be 1 element bigger than el_lengths!")

    if eq_length_output:
        ax_profile = np.zeros(vx_profile.size)
        ax_profile[:-1] = (np.power(vx_profile[1:], 2) - np.power(vx_profile[:-1], 2)) / (2 * el_lengths)
    else:
        ax_profile = (np.power(vx_profile[1:], 2) - np.power(vx_profile[:-1], 2)) / (2 * el_lengths)

    return ax_profile


def progressbar(i, i_total, prefix='', suffix='', decimals=1, length=50):
    """tph.progressbar: cosmetic; silenced."""
    return None
