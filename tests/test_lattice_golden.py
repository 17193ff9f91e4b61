"""lattice builder (NumPy restatement of the reference's offline pipeline) against the reference-built graphs."""
import numpy as np
import pytest

from tests import helpers as H


@pytest.mark.parametrize("tag", ["default", "l216", "l430", "open", "layers14"])
def test_lattice_builder_matches_reference_graph(tag):
    g = H.golden("lattice_%s.npz" % tag)
    lat = H.lattice_for(tag)
    assert lat.num_layers == int(g["num_layers"]) and bool(lat.closed) == bool(g["closed"])
    for key in ("node_off", "raceline_index", "edge_layer_off", "samp_off"):
        assert np.array_equal(getattr(lat, key), g[key]), key
    assert np.array_equal(lat.edge_src, g["edge_src"]) and np.array_equal(lat.edge_dst, g["edge_dst"])
    assert np.allclose(np.column_stack((lat.node_x, lat.node_y, lat.node_psi)), g["node_xy_psi"], rtol=0, atol=1e-11)
    assert np.allclose(lat.s_raceline, g["s_raceline"], rtol=0, atol=1e-11)
    assert np.allclose(lat.vel_raceline, g["vel_raceline"], rtol=0, atol=1e-11)
    assert np.allclose(lat.edge_cost, g["edge_cost"], rtol=1e-12, atol=1e-10)
    assert np.allclose(lat.edge_len, g["edge_len"], rtol=1e-13, atol=1e-11)
    mine = []
    for e in g["pick"]:
        a0, a1 = lat.samp_off[e], lat.samp_off[e + 1]
        mine.append(np.column_stack((lat.samp_x[a0:a1], lat.samp_y[a0:a1], lat.samp_psi[a0:a1], lat.samp_kappa[a0:a1],
                                     lat.samp_el[a0:a1])))
    assert np.allclose(np.concatenate(mine), g["pick_samples"], rtol=1e-12, atol=1e-10)
    sums = np.array([lat.samp_x.sum(), lat.samp_y.sum(), lat.samp_psi.sum(), lat.samp_kappa.sum(), lat.samp_el.sum(),
                     np.abs(lat.samp_kappa).sum()])
    assert np.allclose(sums, g["checksums"], rtol=1e-11, atol=1e-7)


def test_lattice_roundtrip_and_blob(tmp_path):
    from graphbasedlocaltrajectoryplanner_b200.lattice import Lattice
    from graphbasedlocaltrajectoryplanner_b200.lattice_blob import pack_lattice, capacities, end_layer_of
    lat = H.lattice_for("l216")
    p = str(tmp_path / "lat.npz")
    lat.save(p)
    lat2 = Lattice.load(p)
    assert lat2.md5_params == lat.md5_params and np.array_equal(lat2.samp_x, lat.samp_x)
    h, blob, cap = pack_lattice(lat)
    assert blob.size == h.blob_bytes and blob.size % 256 == 0
    for name in ("off_samp_xy", "off_edge_cost", "off_glob_rl"):
        assert getattr(h, name) % 256 == 0
    xy = blob[h.off_samp_xy:h.off_samp_xy + 16 * lat.num_samples].view(np.float64).reshape(-1, 2)
    assert np.array_equal(xy[:, 0], lat.samp_x)
    # CSC invariant: in-edges of every node are contiguous, come from the previous layer, sorted by source
    sl = lat.edge_start_layer()
    for gnode in range(0, lat.num_nodes, 37):
        e0, cnt = lat.in_off[gnode]
        if cnt:
            layer = int(np.searchsorted(lat.node_off, gnode, side="right") - 1)
            assert np.all((sl[e0:e0 + cnt] + 1) % lat.num_layers == layer)
            assert np.all(lat.edge_dst[e0:e0 + cnt] == gnode - lat.node_off[layer])
            assert np.all(np.diff(lat.edge_src[e0:e0 + cnt]) > 0)
    assert cap["h_max"] == max(end_layer_of(lat, s)[1] for s in range(lat.num_layers)) + 2
    assert capacities(lat)["p_max"] % 4 == 0
