#!/usr/bin/env python
"""One-off extraction of *numeric facts* from the reference assets into
``seqdex_b200/robot_data.py``.  Run in the build container only (it reads
/root/reference, which does not exist on the GPU box); the generated module is
committed, so nothing at run time touches the reference tree.

What is extracted (no code, only numbers):
  * the Franka-Panda + Allegro kinematic tree of
    assets/urdf/franka_description/robots/franka_panda_allegro.urdf after
    collapsing fixed joints, in Isaac Gym's DoF order (depth first, siblings
    in lexicographic joint-name order -- SURVEY.md Appendix A.1),
  * one oriented box per collision geometry (URDF <box> verbatim; meshes are
    replaced by the axis-aligned bounding box of the mesh in its own frame),
  * link masses / inertias (URDF <inertial> where present, otherwise box
    volume x 1000 kg/m^3, Isaac Gym's default density),
  * the brick bounding boxes of assets/urdf/blender/origin_obj/*/*.stl.
"""
import os
import re
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

REF = "/root/reference/assets/urdf"
URDF = os.path.join(REF, "franka_description/robots/franka_panda_allegro.urdf")


def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def T(xyz, rpy):
    m = np.eye(4)
    m[:3, :3] = rpy_to_mat(rpy)
    m[:3, 3] = xyz
    return m


def parse_origin(el):
    o = el.find("origin") if el is not None else None
    if o is None:
        return np.eye(4)
    xyz = [float(v) for v in o.get("xyz", "0 0 0").split()]
    rpy = [float(v) for v in o.get("rpy", "0 0 0").split()]
    return T(xyz, rpy)


def mesh_bounds(path, scale):
    if path.lower().endswith(".obj"):
        v = np.array([[float(x) for x in l.split()[1:4]] for l in open(path) if l.startswith("v ")])
    else:
        d = open(path, "rb").read()
        n = struct.unpack("<I", d[80:84])[0]
        if 84 + n * 50 == len(d):
            a = np.frombuffer(d[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n)
            v = a["v"].reshape(-1, 3).astype(np.float64)
        else:
            v = np.array([[float(x) for x in m.groups()] for m in
                          re.finditer(r"vertex\s+(\S+)\s+(\S+)\s+(\S+)", d.decode("latin1"))])
    v = v * np.asarray(scale)
    return v.min(0), v.max(0)


def resolve_mesh(fn):
    if fn.startswith("package://franka_description/"):
        return os.path.join(REF, "franka_description", fn[len("package://franka_description/"):])
    return os.path.normpath(os.path.join(os.path.dirname(URDF), fn))


def main():
    root = ET.parse(URDF).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children = {}
    for j in joints:
        p = j.find("parent").get("link")
        children.setdefault(p, []).append(j)
    child_names = {j.find("child").get("link") for j in joints}
    base = [n for n in links if n not in child_names][0]

    bodies = []   # dict(name, parent, T_parent_joint, axis, lo, hi, shapes, mass parts)
    dofs = []

    def link_payload(name, T_body_link):
        """collision boxes + inertial contributions of URDF link `name` expressed in body frame"""
        shapes, inert = [], []
        l = links[name]
        for c in l.findall("collision"):
            Tc = T_body_link @ parse_origin(c)
            g = c.find("geometry")
            if g.find("box") is not None:
                size = np.array([float(v) for v in g.find("box").get("size").split()])
                shapes.append((Tc, size / 2, name))
            elif g.find("mesh") is not None:
                m = g.find("mesh")
                sc = [float(v) for v in m.get("scale", "1 1 1").split()]
                lo, hi = mesh_bounds(resolve_mesh(m.get("filename")), sc)
                Tm = Tc @ T((lo + hi) / 2, (0, 0, 0))
                shapes.append((Tm, (hi - lo) / 2, name))
        i = l.find("inertial")
        if i is not None and i.find("mass") is not None:
            Ti = T_body_link @ parse_origin(i)
            mass = float(i.find("mass").get("value"))
            it = i.find("inertia")
            I = np.zeros((3, 3))
            if it is not None:
                g = lambda k: float(it.get(k, "0"))
                I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            inert.append((Ti, mass, I))
        else:
            for (Tc, half, _) in shapes:
                vol = 8 * half[0] * half[1] * half[2]
                mass = 1000.0 * vol
                I = np.diag([mass / 3 * (half[1] ** 2 + half[2] ** 2), mass / 3 * (half[0] ** 2 + half[2] ** 2),
                             mass / 3 * (half[0] ** 2 + half[1] ** 2)])
                inert.append((Tc, mass, I))
        return shapes, inert

    def build(name, parent_body, T_parent_joint, joint):
        idx = len(bodies)
        b = dict(name=name, parent=parent_body, T=T_parent_joint, joint=joint, shapes=[], inert=[])
        bodies.append(b)
        if joint is not None:
            dofs.append(idx)

        def absorb(lname, T_body_link):
            s, i = link_payload(lname, T_body_link)
            b["shapes"] += s
            b["inert"] += i
            for j in sorted(children.get(lname, []), key=lambda j: j.get("name")):
                if j.get("type") == "fixed":
                    absorb(j.find("child").get("link"), T_body_link @ parse_origin(j))
        absorb(name, np.eye(4))

        # movable children, depth first, lexicographic joint-name order over the whole merged body
        movable = []

        def collect(lname, T_body_link):
            for j in children.get(lname, []):
                if j.get("type") == "fixed":
                    collect(j.find("child").get("link"), T_body_link @ parse_origin(j))
                else:
                    movable.append((j.get("name"), j, T_body_link @ parse_origin(j)))
        collect(name, np.eye(4))
        for _, j, Tj in sorted(movable, key=lambda t: t[0]):
            build(j.find("child").get("link"), idx, Tj, j)

    build(base, -1, np.eye(4), None)

    out = []
    out.append('"""GENERATED by tools/extract_scene.py from the reference assets (numbers only).\n'
               'Franka-Panda + Allegro tree after collapse_fixed_joints, Isaac Gym DoF order;\n'
               'see SURVEY.md Appendix A.1.  Do not edit by hand."""\n')
    out.append("BODY_NAMES = %r\n" % [b["name"] for b in bodies])
    out.append("BODY_PARENT = %r\n" % [b["parent"] for b in bodies])
    jo, ja, jl = [], [], []
    for b in bodies:
        if b["joint"] is None:
            continue
        Tm = b["T"]
        ax = [float(v) for v in b["joint"].find("axis").get("xyz").split()]
        lim = b["joint"].find("limit")
        jo.append([float(x) for x in Tm[:3, 3]] + [float(x) for x in Tm[:3, :3].reshape(-1)])
        ja.append(ax)
        jl.append([float(lim.get("lower")), float(lim.get("upper"))])
    out.append("# per DoF (child body = dof+1): parent->joint frame [xyz(3), R row-major(9)]\n")
    out.append("JOINT_FRAME = %r\n" % jo)
    out.append("JOINT_AXIS = %r\n" % ja)
    out.append("JOINT_LIMITS = %r\n" % jl)
    out.append("JOINT_NAMES = %r\n" % [b["joint"].get("name") for b in bodies if b["joint"] is not None])
    shapes = []
    for bi, b in enumerate(bodies):
        for (Tc, half, src) in b["shapes"]:
            shapes.append([bi] + [float(x) for x in Tc[:3, 3]] + [float(x) for x in Tc[:3, :3].reshape(-1)] +
                          [float(x) for x in half] + [src])
    out.append("# collision boxes: [body, center xyz(3), R row-major(9), half extents(3), source link]\n")
    out.append("SHAPES = %r\n" % shapes)
    # mass properties per body about body origin
    mp = []
    for b in bodies:
        m_tot, c_tot, I_tot = 0.0, np.zeros(3), np.zeros((3, 3))
        for (Ti, m, I) in b["inert"]:
            m_tot += m
            c_tot += m * Ti[:3, 3]
        c = c_tot / max(m_tot, 1e-12)
        for (Ti, m, I) in b["inert"]:
            R = Ti[:3, :3]
            d = Ti[:3, 3] - c
            I_tot += R @ I @ R.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        mp.append([float(m_tot)] + [float(x) for x in c] + [float(x) for x in I_tot.reshape(-1)])
    out.append("# per body: [mass, com xyz(3) in body frame, inertia about com row-major(9)]\n")
    out.append("BODY_MASSPROPS = %r\n" % mp)

    # bricks
    bricks = {}
    for d in sorted(os.listdir(os.path.join(REF, "blender/origin_obj"))):
        p = os.path.join(REF, "blender/origin_obj", d, d + ".stl")
        if os.path.exists(p):
            lo, hi = mesh_bounds(p, (0.01, 0.01, 0.01))
            bricks[d] = [[round(float(x), 6) for x in lo], [round(float(x), 6) for x in hi]]
    out.append("# brick mesh AABBs (m) in the URDF link frame, mesh scale 0.01 (GS:731)\n")
    out.append("BRICK_AABB = %r\n" % bricks)
    dst = os.path.join(os.path.dirname(__file__), "..", "seqdex_b200", "robot_data.py")
    with open(dst, "w") as f:
        f.write("".join(out))
    print("wrote", dst, "bodies", len(bodies), "dofs", len(dofs), "shapes", len(shapes))
    for i, b in enumerate(bodies):
        print(i, b["name"], "parent", b["parent"], "nshapes", len(b["shapes"]), "mass %.4f" % mp[i][0])


if __name__ == "__main__":
    main()
