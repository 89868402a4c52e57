"""What the edge-edge contacts buy, measured on the CPU oracle (the kernel is bit-identical to it): the lattice of 72
bricks (GS:737-742) is dropped into the box and left to settle with sleeping off; afterwards the TRUE overlap of every pair of bricks is
computed here with the full separating-axis test (15 axes, numpy) -- independent of which contacts the contact step generated.

    python tools/edge_contact_audit.py [envs] [steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from seqdex_b200.scene import Scene  # noqa: E402


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def sat_overlap(ca, Ra, ha, cb, Rb, hb):
    """least overlap of two boxes over the 15 separating axes (negative: separated) and whether an edge pair is that axis"""
    d = cb - ca
    best, edge = np.inf, False
    axes = [(Ra[:, i], False) for i in range(3)] + [(Rb[:, i], False) for i in range(3)]
    for i in range(3):
        for j in range(3):
            n = np.cross(Ra[:, i], Rb[:, j])
            ln = np.linalg.norm(n)
            if ln > 0.03:
                axes.append((n / ln, True))
    for n, is_edge in axes:
        ra = np.abs(Ra.T @ n) @ ha
        rb = np.abs(Rb.T @ n) @ hb
        ov = ra + rb - abs(d @ n)
        if ov < best - (1e-6 if is_edge else 0.0):
            best, edge = ov, is_edge
    return best, edge


def true_penetrations(scene, brick, nbr):
    """brick: the contact step's own state [n][13][NB] (box centre = centre of mass, rows 0-2; quaternion xyzw, rows 3-6)"""
    rows = np.transpose(brick, (0, 2, 1))
    half = np.ctypeslib.as_array(scene.c.br_half).reshape(-1, 3)[:nbr]
    out = []
    for e in range(rows.shape[0]):
        c = rows[e, :nbr, 0:3].astype(np.float64)
        R = [quat_R(rows[e, b, 3:7].astype(np.float64)) for b in range(nbr)]
        rad = np.linalg.norm(half, axis=1)
        for a in range(nbr):
            for b in range(a + 1, nbr):
                if np.linalg.norm(c[a] - c[b]) > rad[a] + rad[b]:
                    continue
                ov, edge = sat_overlap(c[a], R[a], half[a], c[b], R[b], half[b])
                if ov > 0:
                    out.append((ov, edge))
    return np.array(out, dtype=np.float64).reshape(-1, 2)


def settle(edge_contacts, envs, steps, seed=0):
    oracle.build()
    s = Scene(sleep_time=0.0, edge_contacts=edge_contacts)
    nbr = int(s.c.n_bricks)
    e = oracle.OracleEnv(s, envs)
    rows = e.brick_roots()
    rng = np.random.default_rng(seed)
    rows[:, :nbr, 0:2] += rng.uniform(-0.01, 0.01, size=(envs, nbr, 2)).astype(np.float32)
    e.set_brick_roots(rows)
    for _ in range(steps):
        e.simulate()
    return s, e, nbr


def audit(edge_contacts, envs, steps):
    s, e, nbr = settle(edge_contacts, envs, steps)
    pen = true_penetrations(s, e.brick, nbr)
    v = np.linalg.norm(e.brick[:, 7:10, :nbr], axis=1)
    return dict(pairs_overlapping=len(pen), deeper_2mm=int((pen[:, 0] > 2e-3).sum()), deeper_5mm=int((pen[:, 0] > 5e-3).sum()),
                deeper_2mm_edge_axis=int(((pen[:, 0] > 2e-3) & (pen[:, 1] > 0)).sum()), max_mm=float(pen[:, 0].max() * 1e3),
                median_mm=float(np.median(pen[:, 0]) * 1e3), contacts=float(e.ncontact[:, 0].mean()), speed_median=float(np.median(v)),
                speed_p95=float(np.quantile(v, 0.95)))


if __name__ == "__main__":
    envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    for on in (False, True):
        print("edge_contacts", on, audit(on, envs, steps))
