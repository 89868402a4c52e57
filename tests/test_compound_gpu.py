"""Compound free bodies on the GPU: k_simulate<true> == the oracle bit for bit (state, contact counts, sleep counters) with three
hammer-shaped bodies falling into the bin while the hand moves."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_compound_bodies_bit_exact(oracle_lib):
    from seqdex_b200.env import SdxEnv
    from seqdex_b200.scene import Scene
    from tests.test_compound_cpu import HANDLE, HEAD
    s = Scene()
    s.set_free_bodies([{"boxes": [HANDLE, HEAD], "root": [0.2 + 0.05 * i, 0.05 + 0.1 * i, 0.75 + 0.06 * i, 0, 0, np.sin(0.3 * i), np.cos(0.3 * i)] + [0] * 6}
                       for i in range(4)] + [{"boxes": [(0, 0, 0, 0.03, 0.015, 0.0287)], "root": [0.3, 0.25, 0.9, 0, 0, 0, 1] + [0] * 6}])
    n = 5
    g, o = SdxEnv(s, n), oracle_lib.OracleEnv(s, n)
    rng = np.random.default_rng(3)
    seen = 0
    for t in range(90):
        tg = o.dof[:, 0, :].copy()
        tg[:, :23] += rng.uniform(-0.05, 0.05, size=(n, 23)).astype(np.float32)
        tg[:, :23] = np.clip(tg[:, :23], s.dof_lo, s.dof_hi)
        o.dof[:, 2, :] = tg
        g.tensor("DOF")[:, 2, :].copy_(torch.from_numpy(tg))
        g.simulate(); o.simulate()
        seen = max(seen, int(o.ncontact[:, 0].max()))
        if t % 10 == 9:
            torch.cuda.synchronize()
            for name, ov in (("BRICK", o.brick), ("DOF", o.dof), ("LINK", o.link), ("NETF", o.netf), ("NCONTACT", o.ncontact), ("SLEEP", o.slp), ("WSN", o.wsn)):
                gv = g.tensor(name).cpu().numpy()
                assert np.array_equal(gv, ov), (t, name, np.abs(gv.astype(np.float64) - ov.astype(np.float64)).max())
    assert seen > 20 and np.isfinite(o.brick).all()
