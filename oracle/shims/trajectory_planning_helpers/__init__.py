"""
TEST INFRASTRUCTURE -- shim that lets the UNMODIFIED reference Python (/root/reference/graph_ltpl) import
`trajectory_planning_helpers` in this container (package pinned ==0.75 in /root/reference/requirements.txt:5, absent and
not installable offline).  Every `tph.<mod>.<func>` the reference touches is mapped onto oracle/tph_port.py.
Used only by oracle/gen_golden.py (golden-vector generation) -- never by the product.
"""
import types as _types

from oracle import tph_port as _p

for _name in ("calc_splines", "interp_splines", "calc_spline_lengths", "calc_head_curv_an", "calc_head_curv_num",
              "normalize_psi", "calc_vel_profile", "calc_vel_profile_brake", "conv_filt", "calc_ax_profile",
              "progressbar"):
    _ns = _types.SimpleNamespace()
    setattr(_ns, _name, getattr(_p, _name))
    globals()[_name] = _ns

# tph.calc_vel_profile.calc_ax_poss is public in the upstream module
calc_vel_profile.calc_ax_poss = _p.calc_ax_poss  # noqa: F821
