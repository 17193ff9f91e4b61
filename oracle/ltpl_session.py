"""
TEST INFRASTRUCTURE -- stateful (multi-tick) restatement of the reference's iterative memory on top of OracleLTPL.

``OracleLTPL.tick`` restates the FIRST tick after ``set_startpos``.  ``OracleSession`` restates what the reference keeps
between ticks (OnlineTrajectoryHandler.py:64-87) and how the next tick uses it:

    calc_paths        OTH:289-516   start node / constant segment from the last executed trajectory and the (wall-clock)
                                    calculation time, cost reduction along the last solution (GLNT:155-162, GB:478-512),
                                    stitching with the constant part
    get_ref_idx       OTH:518-601   cut index on the last trajectory from the position estimate, delay compensation
    calc_vel_profile  OTH:603-1040  trimming of the memory, profiles behind ``vel_course``, recursive infeasibility
                                    fallback on the backup plan (OTH:950-1006), emergency profile
    export cut        LTPL:401-406  NOTE: ``Graph_LTPL.calc_vel_profile`` cuts the trajectories in the dict object it
                                    shares with the handler, so the memory ``__last_bp_action_set`` holds the CUT rows

The only non-reproducible input of the reference, ``time.time()`` (OTH:353-354, 395), is injected as ``clock``.
This is the groundwork of SURVEY 8(f) rank 1; the CUDA path does not implement it yet.
"""
import time

import numpy as np

from oracle import tph_port as tph
from oracle.ltpl_oracle import ACTION_ID_MAP, OracleLTPL, closest_path_index, get_s_coord

CALC_TIME_SAFETY = 2.0     # ltpl_config_online.ini:91
CALC_TIME_BUFFER_LEN = 5   # ltpl_config_online.ini:94


class OracleSession(object):
    def __init__(self, orc: OracleLTPL, clock=None):
        self.orc = orc
        self.clock = clock or time.time
        self.calc_buffer = []                                  # OTH:62
        self.traj_base_id = 0                                  # OTH:65-75
        self.start_node = None
        self.m_nodes = None                                    # __last_action_set_nodes      {action: [list]}
        self.m_node_idx = None                                 # __last_action_set_node_idx
        self.m_coeff = None                                    # __last_action_set_coeff
        self.m_path = None                                     # __last_action_set_path_param
        self.m_gg = None                                       # __last_action_set_path_gg
        self.m_red = None                                      # __last_action_set_red_len
        self.m_bp = None                                       # __last_bp_action_set
        self.last_path_timestamp = None
        self.last_cut_idx = 0
        self.pos_est = None
        self.em_base_id = None
        self.backup = None                                     # dict(nodes, node_idx, coeff, path, gg) or None
        self.action_id_forced = None
        self.v_start = 0.0
        self.closest_obj_index = None
        self.obj_veh = []
        self.zone_nodes = None                                 # processed zone: (id, set of (layer, node))
        self.prev_action_id = None

    # -- Graph_LTPL.set_startpos -> OTH.set_initial_pose (OTH:181-270) ---------------------------------------------------------
    def set_startpos(self, pos, heading, vel=0.0):
        st = self.orc.set_startpos(np.asarray(pos, dtype=np.float64), float(heading), float(vel))
        self.v_start = float(vel)
        if not (st['in_track'] and st['cor_heading']):
            return True
        self.start_node = st['start_node']
        self.action_id_forced = "straight"
        self.m_coeff = {"straight": [st['coeff']]}
        self.m_path = {"straight": [st['path_param']]}
        self.m_nodes = {"straight": [st['nodes']]}
        self.m_node_idx = {"straight": [st['node_idx']]}
        return False

    # -- Graph_LTPL.calc_paths (LTPL:300-340) -> OTH.calc_paths (OTH:289-516) ----------------------------------------------------
    def calc_paths(self, prev_action_id, object_list=None, blocked_zones=None, prev_traj_idx=0):
        orc, lt = self.orc, self.orc.lat
        self.prev_action_id = prev_action_id
        self.obj_veh = orc.process_object_list(object_list)
        self.closest_obj_index = None                         # OTH.update_objects (OTH:286-288)
        i_sel = prev_traj_idx
        action_id_sel = prev_action_id
        if action_id_sel == 'emergency':
            action_id_sel = self.em_base_id
        if self.action_id_forced is not None:
            action_id_sel = self.action_id_forced
            self.action_id_forced = None

        const_path_seg_exists = self.m_path is not None and action_id_sel in self.m_path
        planned_once = self.last_path_timestamp is not None
        valid_last = (planned_once and const_path_seg_exists and self.m_bp[action_id_sel][i_sel].shape[0] > 2)

        if valid_last:                                         # OTH:325-344 backup plan
            tmp = "follow" if "follow" in self.m_nodes else "straight"
            self.backup = dict(coeff=self.m_coeff[tmp][0], node_idx=self.m_node_idx[tmp][0], nodes=self.m_nodes[tmp][0],
                               path=self.m_path[tmp][0], gg=self.m_gg[tmp][0])
        else:
            self.backup = None

        last_solution_nodes = None
        if planned_once and valid_last:                        # OTH:351-392
            now = self.clock()
            calc_time = now - self.last_path_timestamp
            self.last_path_timestamp = self.clock()
            if len(self.calc_buffer) >= CALC_TIME_BUFFER_LEN:
                self.calc_buffer.pop(0)
            self.calc_buffer.append(calc_time)
            calc_time_avg = float(np.sum(self.calc_buffer) / len(self.calc_buffer))
            bp = self.m_bp[action_id_sel][i_sel]
            s_past = np.diff(bp[1:, 0])
            v_past = bp[1:-1, 5]
            t_approx = np.divide(s_past, v_past, out=np.full(v_past.shape[0], np.inf), where=v_past != 0)   # q11
            t_const = min(calc_time_avg * CALC_TIME_SAFETY, 0.5)
            next_idx = (np.cumsum(t_approx) <= t_const).argmin() + 1
            last_node_idx = self.m_node_idx[action_id_sel][i_sel]
            node_coords = self.m_path[action_id_sel][i_sel][last_node_idx, 0:2]
            predicted_pos = bp[next_idx, 1:3]
            start_node_idx = get_s_coord(node_coords, predicted_pos, only_index=True)[1][1]
            loc_path_start_idx = self.m_node_idx[action_id_sel][i_sel][start_node_idx]
            self.start_node = self.m_nodes[action_id_sel][i_sel][start_node_idx]
            last_solution_nodes = self.m_nodes[action_id_sel][i_sel][start_node_idx:]
        else:                                                  # OTH:393-407
            self.last_path_timestamp = self.clock()
            if const_path_seg_exists and self.start_node in self.m_nodes[action_id_sel][i_sel]:
                start_node_pos = orc.node_pos(self.start_node[0], self.start_node[1])[0]
                loc_path_start_idx = closest_path_index(self.m_path[action_id_sel][i_sel][:, 0:2], start_node_pos)[0]
                start_node_idx = self.m_nodes[action_id_sel][i_sel].index(self.start_node)
            else:
                loc_path_start_idx = 0
                start_node_idx = 0

        const_path_seg = None
        if const_path_seg_exists:
            const_path_seg = self.m_path[action_id_sel][i_sel][:loc_path_start_idx + 1, :]

        # zones: a new zone is processed with the CURRENT start node (GLNT:43-99), afterwards its node set is fixed
        zone = set()
        if blocked_zones:
            if len(blocked_zones) != 1:
                raise NotImplementedError("more than one blocked zone per tick is not a defined input of the reference")
            zid = list(blocked_zones.keys())[0]
            if self.zone_nodes is None or self.zone_nodes[0] != zid:
                # a zone under a new id: update_zone (OLI:155-237) flags the old one as removed; GLNT:78-91 keeps its nodes
                # within the next BLOCK_N_LAYERS_WHEN_REMOVING_ZONE = 0 layers, i.e. none -- it is empty from this tick on
                # and dropped by the next update_zone; the new zone is processed with the CURRENT start node
                self.zone_nodes = (zid, orc.zone_removed_nodes(self.start_node, blocked_zones))
            zone = self.zone_nodes[1]
        elif self.zone_nodes is not None:
            zone = self.zone_nodes[1]                          # Graph_LTPL keeps __obj_zone when no dict is passed

        cost_factor = None                                     # GLNT:155-162 + GB:478-512 (planning_range copy)
        if last_solution_nodes is not None:
            cost_factor = {}
            w = self.orc.p['w_last_edges']
            for i in range(min(len(last_solution_nodes) - 1, len(w))):
                a, b = last_solution_nodes[i], last_solution_nodes[i + 1]
                if a[0] is None:
                    continue
                if (a[0], a[1]) in zone or (b[0], b[1]) in zone:
                    continue                                   # vertex absent in the filtered graph -> ValueError -> pass
                try:
                    e = lt.edge_id(a[0], a[1], b[1])
                except (KeyError, ValueError):
                    continue
                cost_factor[e] = cost_factor.get(e, 1.0) * w[i]

        res, self.closest_obj_index = orc.main_online_path_gen(self.start_node, self.obj_veh, action_id_sel,
                                                               const_path_seg, self.pos_est, zone=zone,
                                                               cost_factor=cost_factor)

        last_pp = self.m_path[action_id_sel][i_sel] if const_path_seg_exists else None
        for name in list(res['nodes'].keys()):                 # OTH:433-472
            if not const_path_seg_exists:
                continue
            pp = res['path_param'][name][0]
            if loc_path_start_idx > 0:
                pp = np.concatenate((last_pp[:loc_path_start_idx, :], pp))
                if np.size(last_pp, axis=0) == loc_path_start_idx:
                    j = loc_path_start_idx - 1
                    pp[j, 4] = np.sqrt(np.power(np.diff(pp[j:j + 2, 0]), 2) + np.power(np.diff(pp[j:j + 2, 1]), 2))
                res['path_param'][name][0] = pp
            res['node_idx'][name][0] = np.concatenate((np.array(self.m_node_idx[action_id_sel][i_sel][:start_node_idx]),
                                                       np.array(res['node_idx'][name][0]) + loc_path_start_idx))
            if start_node_idx > 0:
                res['nodes'][name][0] = np.concatenate((self.m_nodes[action_id_sel][i_sel][:start_node_idx],
                                                        res['nodes'][name][0])).tolist()
                res['coeff'][name][0] = np.concatenate((self.m_coeff[action_id_sel][i_sel][:start_node_idx],
                                                        res['coeff'][name][0]))

        if not res['nodes']:                                   # OTH:475-506 "track blocked"
            if const_path_seg_exists and const_path_seg.shape[0] > 2:
                loc_path_start_idx += 1
                start_node_idx += 1
                res['path_param'][action_id_sel] = [last_pp[:loc_path_start_idx, :]]
                res['node_idx'][action_id_sel] = [np.array(self.m_node_idx[action_id_sel][i_sel][:start_node_idx])]
                res['nodes'][action_id_sel] = [self.m_nodes[action_id_sel][i_sel][:start_node_idx]]
                res['coeff'][action_id_sel] = [self.m_coeff[action_id_sel][i_sel][:start_node_idx]]
                res['red_len'][action_id_sel] = [True]

        self.m_nodes = res['nodes']                            # OTH:508-513
        self.m_node_idx = res['node_idx']
        self.m_coeff = res['coeff']
        self.m_path = res['path_param']
        self.m_red = res['red_len']
        self.tie = res.get('tie', {})
        return {k: [v[0].copy()] for k, v in self.m_path.items()}

    # -- OTH.get_ref_idx (OTH:518-601) ---------------------------------------------------------------------------------------
    def get_ref_idx(self, action_id_sel, i_sel, pos_est):
        self.pos_est = pos_est
        planned_once = self.m_bp is not None
        valid_last = (planned_once and action_id_sel in self.m_bp and np.size(self.m_bp[action_id_sel][i_sel], axis=0) > 0)
        valid_this = len(list(self.m_node_idx.keys())) > 0
        if planned_once and valid_last:
            bp = self.m_bp[action_id_sel][i_sel]
            idx_nb = get_s_coord(bp[:, 1:3], pos_est, bp[:, 0], only_index=True)[1]
            cut_index = idx_nb[0]
            s_past = np.diff(bp[cut_index:, 0])
            v_past = bp[cut_index:-1, 5]
            t_approx = np.divide(s_past, v_past, out=np.full(v_past.shape[0], np.inf), where=v_past != 0)
            vel_idx = min((np.cumsum(t_approx) <= self.orc.p['delaycomp']).argmin() + 1, v_past.shape[0] - 1)
            vel_plan = bp[cut_index + vel_idx, 5]
            vel_course = bp[cut_index:cut_index + vel_idx, 5]
            cut_index_pos = self.last_cut_idx + cut_index
            if valid_this:
                tmp = list(self.m_node_idx.keys())[0]
                cut_layer = max(np.argmin(np.array(self.m_node_idx[tmp][0]) < cut_index_pos) - 2, 0)
                cut_index_layer = self.m_node_idx[tmp][0][cut_layer]
            else:
                cut_layer = 0
                cut_index_layer = 0
        else:
            cut_index_pos = 0
            cut_layer = 0
            cut_index_layer = 0
            vel_course = np.array([])
            vel_plan = self.v_start
        self.last_cut_idx = cut_index_pos - cut_index_layer
        return int(cut_index_pos), int(cut_layer), float(vel_plan), vel_course

    # -- Graph_LTPL.calc_vel_profile (LTPL:344-408) -> OTH.calc_vel_profile (OTH:603-1040) --------------------------------------
    def calc_vel_profile(self, pos_est, vel_est, vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0),
                         ax_max_machines=np.atleast_2d([100.0, 5.0]), safety_d=30.0, incl_emerg_traj=False):
        orc = self.orc
        pos_est = np.asarray(pos_est, dtype=np.float64)
        cut_index_pos, cut_layer, vel_plan, vel_course = self.get_ref_idx(self.prev_action_id, 0, pos_est)
        if type(local_gg) is not dict and (type(local_gg) is not tuple or len(local_gg) != 2):   # OTH:649-653
            raise ValueError("Provided local_gg does not satisfy requested format! Read parameter documentation.")
        vk = dict(vel_max=vel_max, gg_scale=gg_scale, ax_max_machines=np.asarray(ax_max_machines, dtype=np.float64))
        self.traj_base_id += 10
        if orc.old_gg_scale is None:
            orc.old_gg_scale = gg_scale                        # VPFB.update_dyn_parameters (VPFB:60, 80-81)
        self.m_bp = {}
        self.m_gg = {}
        gg_vel = {}
        ids = {}
        for action_id in list(self.m_path.keys()):
            self.m_bp[action_id] = []
            self.m_gg[action_id] = []
            ids[action_id] = self.traj_base_id + ACTION_ID_MAP.get(action_id, 9)
            full = self.m_path[action_id][0]
            if type(local_gg) is dict:                     # location dependent friction, aligned with the path (OTH:708)
                gg_full = np.asarray(local_gg[action_id][0], dtype=np.float64)
            else:
                gg_full = np.ones((full.shape[0], 2)) * tuple(local_gg)
            pp = full[cut_index_pos:, :]                       # OTH:700-706
            gg = gg_full[cut_index_pos:, :]
            gg_vel[action_id] = gg
            cut_index_layer = self.m_node_idx[action_id][0][cut_layer]
            self.m_node_idx[action_id][0] = np.array(self.m_node_idx[action_id][0][cut_layer:]) - cut_index_layer
            self.m_path[action_id][0] = full[cut_index_layer:, :]
            self.m_gg[action_id].append(gg_full[cut_index_layer:, :])
            self.m_coeff[action_id][0] = self.m_coeff[action_id][0][cut_layer:, :]
            self.m_nodes[action_id][0] = self.m_nodes[action_id][0][cut_layer:]

            bp_out, vel_bound = orc.vel_one(action_id, pp, gg, self.m_nodes[action_id][0][-1], self.m_red[action_id][0],
                                            vel_plan, vel_course, vel_est, self.closest_obj_index, self.obj_veh,
                                            self.pos_est, safety_d, vk)

            if vel_bound or action_id in ["follow", "straight"]:   # OTH:943-1025
                if vel_bound or self.backup is None:
                    self.m_bp[action_id].append(bp_out)
                else:                                          # recursive infeasibility: brake on the backup path
                    bk = self.backup
                    self.m_node_idx[action_id][0] = np.array(bk['node_idx'][cut_layer:]) - cut_index_layer
                    self.m_path[action_id][0] = bk['path'][cut_index_layer:, :]
                    self.m_gg[action_id][0] = bk['gg'][cut_index_layer:, :]
                    self.m_coeff[action_id][0] = bk['coeff'][cut_layer:, :]
                    self.m_nodes[action_id][0] = bk['nodes'][cut_layer:]
                    i = vel_course.shape[0]
                    vx = tph.calc_vel_profile_brake(loc_gg=bk['gg'][cut_index_pos + i:, :],   # VPFB:229-255, no gg scale
                                                    kappa=bk['path'][cut_index_pos + i:, 3],
                                                    el_lengths=bk['path'][(cut_index_pos + i):-1, 4], v_start=vel_plan,
                                                    dyn_model_exp=orc.dyn_model_exp, drag_coeff=orc.drag_coeff,
                                                    m_veh=orc.m_veh)
                    vx = np.concatenate((vel_course, vx))
                    vx_f = tph.conv_filt(signal=vx, filt_window=orc.p['filt_window_width'], closed=False)
                    ax_f = tph.calc_ax_profile(vx_profile=vx_f, el_lengths=bk['path'][cut_index_pos:-1, 4])
                    ax_f[np.logical_and(np.isclose(vx_f[:-1], 0.0), np.isclose(ax_f, 0.0))] = -5.0
                    s = np.concatenate(([0], np.cumsum(bk['path'][cut_index_pos:-1, 4])))
                    self.m_bp[action_id].append(np.column_stack((s, bk['path'][cut_index_pos:, 0:4], vx_f,
                                                                 np.append(ax_f, [0.0]))))
            else:                                              # action set removed
                self.m_coeff[action_id][0] = []
                self.m_path[action_id][0] = []
                self.m_gg[action_id][0] = []
                self.m_nodes[action_id][0] = []
                self.m_node_idx[action_id][0] = []
            if not any([bool(np.size(t)) for t in self.m_nodes[action_id]]):   # OTH:1016-1025
                for d in (self.m_coeff, self.m_path, self.m_gg, self.m_nodes, self.m_node_idx, self.m_red, self.m_bp):
                    d.pop(action_id)
                ids.pop(action_id)
                gg_vel.pop(action_id)

        if incl_emerg_traj and self.m_bp:                      # OTH:1027-1034
            self.em_base_id = list(self.m_bp.keys())[0]
            traj = self.m_bp[self.em_base_id][0]
            el = np.diff(traj[:, 0])
            v_brake = tph.calc_vel_profile_brake(kappa=traj[:, 4], el_lengths=el, v_start=traj[0, 5], drag_coeff=0.854,
                                                 m_veh=1160.0, loc_gg=gg_vel[self.em_base_id])
            a_brake = tph.calc_ax_profile(vx_profile=v_brake, el_lengths=el[:len(v_brake)], eq_length_output=True)
            self.m_bp['emergency'] = [np.column_stack((traj[:len(v_brake), 0:5], v_brake, a_brake))]
            ids['emergency'] = ids[self.em_base_id]

        n_exp = orc.p['nmbr_export_points']                    # LTPL:401-406: cuts the SHARED dict -> memory is cut too
        for action_id in self.m_bp:
            self.m_bp[action_id][0] = self.m_bp[action_id][0][:n_exp, :]
        return {k: [v[0].copy()] for k, v in self.m_bp.items()}, dict(ids)
