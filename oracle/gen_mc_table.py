"""Generate the marching-cubes case table used by BOTH the CPU oracle (oracle/mc_oracle.py) and the CUDA
post-pass (disn_b200/csrc/mc_table.h).  TEST/BUILD INFRASTRUCTURE: run once, outputs are committed.

The reference's mesher is a closed-source binary (isosurface/computeMarchingCubes, Vega FEM), so its
case table is unavailable: topology parity against it is UNPINNED.  This table is derived, not recalled:
for each of the 256 sign configurations the iso-polygons are traced face by face with one fixed rule
(on every cube face, walking the face boundary counter-clockwise as seen from outside the cube, a
segment runs from each crossing that ENTERS the inside region to the next crossing that LEAVES it,
i.e. ambiguous faces separate their inside corners).  The rule depends only on the four corner signs of
a face, so the two cubes sharing a face always agree and the mesh is crack-free.  Loops are fan
triangulated from their lowest edge id, wound so that normals point towards the outside (positive side).

Conventions:  corner i at (x,y,z) = (i&1, (i>>1)&1, (i>>2)&1);  bit i of the case index is set when
value[corner i] < iso ("inside");  edge e joins EDGE_CORNERS[e] and is stored on the lattice as
(axis, dx, dy, dz) = EDGE_LATTICE[e]: the edge along `axis` whose lower corner is cell + (dx,dy,dz).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# 12 edges: 4 along x, 4 along y, 4 along z
EDGE_LATTICE = []   # (axis, dx, dy, dz)
for axis in range(3):
    for b in range(2):
        for a in range(2):
            d = [0, 0, 0]
            others = [k for k in range(3) if k != axis]
            d[others[0]] = a
            d[others[1]] = b
            EDGE_LATTICE.append((axis, d[0], d[1], d[2]))


def corner_id(x, y, z):
    return x | (y << 1) | (z << 2)


EDGE_CORNERS = []
for axis, dx, dy, dz in EDGE_LATTICE:
    lo = [dx, dy, dz]
    hi = list(lo)
    hi[axis] += 1
    EDGE_CORNERS.append((corner_id(*lo), corner_id(*hi)))
EDGE_OF = {frozenset(c): e for e, c in enumerate(EDGE_CORNERS)}

# faces: 4 corners counter-clockwise seen from OUTSIDE the cube
FACES = []
for axis in range(3):
    u, v = [k for k in range(3) if k != axis]      # u x v = +/- axis
    sign = 1 if (axis, u, v) in ((0, 1, 2), (1, 2, 0), (2, 0, 1)) else -1
    for side in range(2):
        quad = []
        for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):     # CCW seen from +axis when u x v = +axis
            c = [0, 0, 0]
            c[axis] = side
            c[u] = a
            c[v] = b
            quad.append(corner_id(*c))
        outward_positive = (side == 1)
        ccw_from_positive = (sign == 1)
        if outward_positive != ccw_from_positive:
            quad.reverse()
        FACES.append(quad)


def triangulate_case(case):
    inside = [(case >> i) & 1 for i in range(8)]
    nxt = {}                                   # directed segments: crossing edge -> crossing edge
    for quad in FACES:
        n = 4
        crossings = []                         # (position k on the face boundary, edge id, kind)
        for k in range(n):
            a, b = quad[k], quad[(k + 1) % n]
            if inside[a] != inside[b]:
                crossings.append((k, EDGE_OF[frozenset((a, b))], "enter" if inside[b] else "leave"))
        for idx, (k, e, kind) in enumerate(crossings):
            if kind == "enter":                # next crossing (cyclically) necessarily leaves
                k2, e2, kind2 = crossings[(idx + 1) % len(crossings)]
                assert kind2 == "leave"
                assert e not in nxt
                nxt[e] = e2
    tris = []
    seen = set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        cur = nxt[start]
        while cur != start:
            loop.append(cur)
            seen.add(cur)
            cur = nxt[cur]
        assert len(loop) >= 3
        tris.extend(triangulate_loop(loop))    # loop order already gives normals towards the outside
    return tris


FACE_EDGE_SETS = [frozenset(EDGE_OF[frozenset((q[k], q[(k + 1) % 4]))] for k in range(4)) for q in FACES]


def coplanar(e1, e2):
    return any(e1 in fs and e2 in fs for fs in FACE_EDGE_SETS)


def all_triangulations(idx):
    """all triangulations of the polygon idx[0..n-1] (indices into the loop), as lists of index triples"""
    n = len(idx)
    if n < 3:
        return [[]]
    if n == 3:
        return [[tuple(idx)]]
    out = []
    for k in range(1, n - 1):                  # triangle (idx[0], idx[k], idx[n-1]) splits the polygon
        for left in all_triangulations(idx[:k + 1]):
            for right in all_triangulations(idx[k:]):
                out.append(left + [(idx[0], idx[k], idx[n - 1])] + right)
    return out


def triangulate_loop(loop):
    """Triangulate so that no diagonal lies in a cube face: a diagonal inside a face could coincide with a
    diagonal of the neighbouring cube and make the mesh non-manifold.  Deterministic: first valid
    triangulation in enumeration order (fans from the lowest edge id come first)."""
    n = len(loop)
    cands = all_triangulations(list(range(n)))
    def bad_diagonals(tri_list):
        bad = 0
        for t in tri_list:
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if (b - a) % n in (1, n - 1):
                    continue                    # polygon side (a face segment)
                if coplanar(loop[a], loop[b]):
                    bad += 1
        return bad
    best = min(cands, key=bad_diagonals)        # min() keeps the first minimum
    assert bad_diagonals(best) == 0, loop
    return [(loop[a], loop[b], loop[c]) for a, b, c in sorted(best)]


def build():
    table = -np.ones((256, 16), dtype=np.int8)
    ntri = np.zeros(256, dtype=np.int8)
    for case in range(256):
        tris = triangulate_case(case)
        assert len(tris) <= 5, (case, len(tris))
        ntri[case] = len(tris)
        flat = [e for t in tris for e in t]
        table[case, :len(flat)] = flat
    return table, ntri


def orientation_check(table, ntri):
    """every triangle's normal must point from the inside corners towards the outside corners"""
    pos = np.array([[(i & 1), (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.float64)
    mid = np.array([(pos[a] + pos[b]) / 2 for a, b in EDGE_CORNERS])
    for case in range(1, 255):
        ins = np.array([(case >> i) & 1 for i in range(8)], bool)
        # per connected patch: area vector . sum(outside - inside endpoints) > 0
        tl = [[int(v) for v in table[case, 3 * t:3 * t + 3]] for t in range(ntri[case])]
        patches = []
        for e in tl:
            hit = [p for p in patches if any(set(e) & set(q) for q in p)]
            merged = [e] + [q for p in hit for q in p]
            patches = [p for p in patches if p not in hit] + [merged]
        for fan in patches:
            area = np.zeros(3)
            grad = np.zeros(3)
            edges = set()
            for e in fan:
                p0, p1, p2 = mid[e[0]], mid[e[1]], mid[e[2]]
                area += np.cross(p1 - p0, p2 - p0)
                edges.update(e)
            for k in edges:
                a, b = EDGE_CORNERS[k]
                grad += (pos[b] - pos[a]) if ins[a] else (pos[a] - pos[b])
            assert np.dot(area, grad) > 0, (case, fan)


def main():
    table, ntri = build()
    orientation_check(table, ntri)
    np.savez(os.path.join(HERE, "mc_table.npz"), table=table, ntri=ntri,
             edge_lattice=np.array(EDGE_LATTICE, dtype=np.int8), edge_corners=np.array(EDGE_CORNERS, dtype=np.int8))
    hdr = os.path.join(HERE, "..", "disn_b200", "csrc", "mc_table.h")
    with open(hdr, "w") as f:
        f.write("// GENERATED by oracle/gen_mc_table.py -- do not edit.  See that file for the derivation rule.\n")
        f.write("#pragma once\n#include <stdint.h>\nnamespace disn {\n")
        f.write("// triangles per case\n__constant__ int8_t kMcNumTris[256] = {%s};\n" % ",".join(str(int(v)) for v in ntri))
        f.write("// up to 5 triangles x 3 edge ids per case, -1 terminated\n__constant__ int8_t kMcTris[256][16] = {\n")
        for row in table:
            f.write("  {%s},\n" % ",".join(str(int(v)) for v in row))
        f.write("};\n// edge id -> (axis, dx, dy, dz): lattice edge along `axis` with lower corner cell+(dx,dy,dz)\n")
        f.write("__constant__ int8_t kMcEdge[12][4] = {%s};\n" % ",".join("{%d,%d,%d,%d}" % e for e in EDGE_LATTICE))
        f.write("}  // namespace disn\n")
    print("cases with triangles:", int((ntri > 0).sum()), "max tris:", int(ntri.max()), "total tris:", int(ntri.sum()))


if __name__ == "__main__":
    main()
