"""Compound free bodies (a body = several boxes; sdx_scene_t::n_bshapes / bs_body / bs_c) in the CPU oracle: mass properties from
the boxes, and a T-shaped body (a hammer: handle + head) coming to rest on BOTH its boxes."""
import numpy as np

from seqdex_b200.scene import Scene

HANDLE, HEAD = (0, 0.056, 0, 0.016, 0.165, 0.018), (0, 0.245, 0, 0.0155, 0.0235, 0.0665)      # harmmer.obj x 0.01: 3.2 x 33 x 3.6 cm, 3.1 x 4.7 x 13.3 cm


def _scene(roots):
    s = Scene()
    s.set_free_bodies([{"boxes": [HANDLE, HEAD], "root": r} for r in roots])
    return s


def test_mass_properties_of_a_compound_body():
    s = _scene([[0.25, 0.19, 0.9, 0, 0, 0, 1] + [0] * 6])
    c = s.c
    assert c.n_bricks == 1 and c.n_bshapes == 2 and list(c.bs_body[:2]) == [0, 0]
    m = [567.0 * 8 * b[3] * b[4] * b[5] for b in (HANDLE, HEAD)]
    assert abs(1.0 / c.br_invm[0] - sum(m)) < 1e-6
    com_y = (m[0] * HANDLE[1] + m[1] * HEAD[1]) / sum(m)
    assert abs(c.br_coff[1] - com_y) < 1e-6 and abs(c.br_coff[0]) < 1e-9 and abs(c.br_coff[2]) < 1e-9
    np.testing.assert_allclose([c.bs_c[1], c.bs_c[4]], [HANDLE[1] - com_y, HEAD[1] - com_y], atol=1e-6)   # box centres relative to the COM
    assert 1.0 / c.br_invI[1] < 1.0 / c.br_invI[0]            # spinning about the handle's axis is the easy one


def test_hammer_comes_to_rest_on_handle_and_head(oracle_lib):
    """dropped onto the table beside the bin with the head's long side vertical: it tips over and comes to rest with BOTH boxes on the
    table top -- the handle alone (a single box) would lie 1.8 cm lower at the head's end"""
    s = _scene([[0.6, -0.42, 0.75, 0, 0, 0, 1] + [0] * 6])            # clear of the bin and of the base-plate
    e = oracle_lib.OracleEnv(s, 1)
    for _ in range(240):
        e.simulate()
    rows = e.brick_roots()[0, 0]
    q = rows[3:7]
    R = np.array([[1 - 2 * (q[1] ** 2 + q[2] ** 2), 2 * (q[0] * q[1] - q[2] * q[3]), 2 * (q[0] * q[2] + q[1] * q[3])],
                  [2 * (q[0] * q[1] + q[2] * q[3]), 1 - 2 * (q[0] ** 2 + q[2] ** 2), 2 * (q[1] * q[2] - q[0] * q[3])],
                  [2 * (q[0] * q[2] - q[1] * q[3]), 2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[0] ** 2 + q[1] ** 2)]])

    def lowest(box):
        c, h = np.array(box[:3]), np.array(box[3:])
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h + c
        return (rows[0:3] + corners @ R.T)[:, 2].min()
    assert abs(lowest(HEAD) - 0.6) < 3e-3 and abs(lowest(HANDLE) - 0.6) < 3e-3, (lowest(HEAD), lowest(HANDLE))   # both boxes touch the table top (z = 0.6)
    assert np.abs(rows[7:13]).max() < 1e-3 and e.slp[0, 0] >= s.c.sleep_substeps                                      # at rest and asleep
