"""Procedural UV-sphere template: synthetic INPUT for tests, smoke() and bench.py (the reference's shipped OBJ templates
cannot travel to the GPU box and are not copied).  Same construction as the shipped files — 32 segments, `rings` rings,
u = 1/4 + atan2(x, z) / 2pi, v = 1 - polar / pi, one vt per pole triangle: 482 v / 559 vt / 960 f for 16 rings; the set of
(vertex, uv) pairs equals the shipped 16-ring template's (checked in the authoring container).  Pure Python, no torch."""
import math


def write_uvsphere_obj(path, rings=16, segments=32):
    v, vt, f = [], [], []
    v.append((0.0, 1.0, 0.0))                       # north pole = vertex 0
    for r in range(1, rings):
        pol = math.pi * r / rings
        for s in range(segments):
            th = 2 * math.pi * (s / segments - 0.25)
            v.append((math.sin(th) * math.sin(pol), math.cos(pol), math.cos(th) * math.sin(pol)))
    v.append((0.0, -1.0, 0.0))
    south = len(v) - 1
    vid = lambda r, s: 1 + (r - 1) * segments + (s % segments)
    # vt grid for ring vertices: (segments+1) columns
    tid = {}
    for r in range(1, rings):
        for s in range(segments + 1):
            tid[(r, s)] = len(vt)
            vt.append((s / segments, 1 - r / rings))
    for s in range(segments):                       # pole fans
        tn = len(vt)
        vt.append(((s + 0.5) / segments, 1.0))
        f.append(((0, tn), (vid(1, s), tid[(1, s)]), (vid(1, s + 1), tid[(1, s + 1)])))
        ts = len(vt)
        vt.append(((s + 0.5) / segments, 0.0))
        f.append(((south, ts), (vid(rings - 1, s + 1), tid[(rings - 1, s + 1)]), (vid(rings - 1, s), tid[(rings - 1, s)])))
    for r in range(1, rings - 1):
        for s in range(segments):
            a, b = (vid(r, s), tid[(r, s)]), (vid(r, s + 1), tid[(r, s + 1)])
            c, d = (vid(r + 1, s), tid[(r + 1, s)]), (vid(r + 1, s + 1), tid[(r + 1, s + 1)])
            f.append((a, c, d))
            f.append((a, d, b))
    with open(path, "w") as fh:
        fh.write("# procedural uv sphere (oracle/mesh.py:write_uvsphere_obj)\n")
        for p in v:
            fh.write("v %.6f %.6f %.6f\n" % p)
        for t in vt:
            fh.write("vt %.6f %.6f\n" % t)
        for tri in f:
            fh.write("f " + " ".join(f"{a + 1}/{b + 1}" for a, b in tri) + "\n")
    return path
