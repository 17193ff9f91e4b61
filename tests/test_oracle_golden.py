"""Pins the oracle (oracle/ltpl_oracle.py, oracle/tph_port.py) against golden vectors produced by executing the
reference's own Python files in the build container (oracle/gen_golden.py)."""
import numpy as np
import pytest

from tests import helpers as H

TAGS = ("default", "l216", "l430", "open", "layers14")


def _run(tag, idx=None):
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("ticks_%s.npz" % tag)
    orc = OracleLTPL(H.lattice_for(tag))
    vk = dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=g["ax_max_machines"], safety_d=30.0)
    n = g["sc_pos"].shape[0]
    for b in (range(n) if idx is None else idx):
        rec = orc.tick(g["sc_pos"][b], g["sc_heading"][b], g["sc_vel"][b], H.object_list(g, b), vk)
        H.compare_record(rec, g, b, ctx=tag)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_ticks(tag):
    _run(tag)


def test_oracle_config1_min_example():
    """main_min_example.py start pose + static dummy object; API-default velocity arguments."""
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("config1_min_example.npz")
    orc = OracleLTPL(H.lattice_for("default"))
    x, y, th, v, ln = (float(a) for a in g["obj"])
    obj = [{'id': 1, 'type': 'physical', 'X': x, 'Y': y, 'theta': th, 'length': ln, 'width': 2.5, 'v': v}]
    for b in range(2):
        rec = orc.tick(g["sc_pos"][b], g["sc_heading"][b], g["sc_vel"][b], obj, {})
        H.compare_record(rec, g, b, prefix="", ctx="config1")
    assert int(g["path_len"][0, 0]) > 0 and int(g["path_len"][1, 1]) > 0   # straight / follow exercised


def test_oracle_zones_and_emergency():
    """blocked zones (GLNT:43-99, 'nodes' type, first tick) and the emergency trajectory (OTH:1027-1034)."""
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("ticks_ext_default.npz")
    orc = OracleLTPL(H.lattice_for("default"))
    vk = dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=g["ax_max_machines"], safety_d=30.0,
              incl_emerg_traj=True)
    n = g["sc_pos"].shape[0]
    assert int((g["zone_layers"][:, 0] >= 0).sum()) >= n // 2 and int((g["em_len"] > 0).sum()) >= n // 2
    for b in range(n):
        rec = orc.tick(g["sc_pos"][b], g["sc_heading"][b], g["sc_vel"][b], H.object_list(g, b), vk,
                       blocked_zones=H.zone_of(g, b))
        H.compare_record(rec, g, b, prefix="", ctx="ext")
        H.compare_emergency(rec, g, b, ctx="ext")


@pytest.mark.parametrize("name", sorted(H.VARIANTS))
def test_oracle_parameter_variants(name):
    """PDtan follow controller, friction-ellipse exponents 1.5 / 2.0, other mass / drag, gg scale, asymmetric gg, ego
    velocity estimate != planned velocity -- all against the unmodified reference."""
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("ticks_variants_default.npz")
    sub = H._Sub(g, name)
    online, veh, vel, dv = H.VARIANTS[name]
    orc = OracleLTPL(H.lattice_for("default"), online=online, **veh)
    vk = dict(vel, ax_max_machines=g["ax_max_machines"])
    n = sub["sc_pos"].shape[0]
    for b in range(n):
        rec = orc.tick(sub["sc_pos"][b], sub["sc_heading"][b], sub["sc_vel"][b], H.object_list(sub, b), vk,
                       vel_est=sub["sc_vel"][b] + dv)
        H.compare_record(rec, sub, b, prefix="", ctx=name)


def test_oracle_explicit_predictions():
    """objects with an explicit 'prediction' array (OLI:117-119, GLNT:180-189, quirk q14) against the reference."""
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("ticks_pred_default.npz")
    orc = OracleLTPL(H.lattice_for("default"))
    vk = dict(vel_max=100.0, gg_scale=1.0, local_gg=(5.0, 5.0), ax_max_machines=g["ax_max_machines"], safety_d=30.0)
    assert int((g["sc_n_pred"] >= 0).sum()) > 20
    for b in range(g["sc_pos"].shape[0]):
        rec = orc.tick(g["sc_pos"][b], g["sc_heading"][b], g["sc_vel"][b], H.object_list(g, b), vk)
        H.compare_record(rec, g, b, prefix="", ctx="pred")


def test_oracle_location_dependent_local_gg():
    """calc_vel_profile(local_gg={action: [ndarray(P, 2)]}) (OTH:649-666, VPFB:194-227): friction as a function of the
    position along every path, emergency trajectory on (raw local_gg of its base trajectory, OTH:1030)."""
    from oracle.ltpl_oracle import OracleLTPL
    g = H.golden("ticks_ggpp_default.npz")
    orc = OracleLTPL(H.lattice_for("default"))
    vk = dict(vel_max=100.0, gg_scale=1.0, ax_max_machines=g["ax_max_machines"], safety_d=30.0, incl_emerg_traj=True)
    n = g["sc_pos"].shape[0]
    assert int((g["em_len"] > 0).sum()) >= n // 2 and int((g["full_traj_len"][:, 1] > 0).sum()) >= 4
    for b in range(n):
        rec = orc.tick(g["sc_pos"][b], g["sc_heading"][b], g["sc_vel"][b], H.object_list(g, b), vk,
                       gg_fn=H.local_gg_field)
        H.compare_record(rec, g, b, ctx="ggpp")
        H.compare_emergency(rec, g, b, ctx="ggpp")
