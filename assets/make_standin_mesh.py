#!/usr/bin/env python3
"""Deterministic procedural stand-in for the reference's missing `public_html/bunny2.obj`.

The reference mount lists the bunny as a stripped large blob (`.MISSING_LARGE_BLOBS:1`);
every "bunny" number produced by this repo is on THIS mesh and is labelled "stand-in mesh".
If the real file is supplied, load it with `wpt_load_obj` instead.

The mesh is a displaced icosphere (subdivision n -> 20 * 4**n triangles; n = 6 gives
81 920) written in exactly the dialect `src_ts/client/obj_parser.ts:3-51` accepts:
`v x y z` and `f a/b/c`-free `f a b c` lines, single spaces, triangles only, 1-based.
Raw coordinates are chosen so that the client's `*(8, 8, -8)` (`src_ts/client/index.ts:216-220`)
followed by `*0.5 + (0,0,5)` (`src/wasm_interface.rs:297-313`) puts the blob on the
floor plane y = -1 in front of the bunny-scene camera.

Only + - * / sqrt are used (all IEEE-exact in float64) and coordinates are printed with
a fixed 6-decimal format, so the file is bit-identical on every machine.
"""
import sys
import numpy as np


def icosphere(subdiv):
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.sqrt((v * v).sum(1))[:, None]
    for _ in range(subdiv):
        nv = len(v)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        es = np.sort(e, 1)
        key = es[:, 0] * nv + es[:, 1]
        uk, inv = np.unique(key, return_inverse=True)
        a, b = uk // nv, uk % nv
        mid = v[a] + v[b]
        mid /= np.sqrt((mid * mid).sum(1))[:, None]
        v = np.concatenate([v, mid], 0)
        m = inv.reshape(3, -1).T + nv      # midpoints of edges 01, 12, 20 per face
        f = np.concatenate([np.stack([f[:, 0], m[:, 0], m[:, 2]], 1), np.stack([f[:, 1], m[:, 1], m[:, 0]], 1),
                            np.stack([f[:, 2], m[:, 2], m[:, 1]], 1), np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)], 0)
    return v, f


def standin(subdiv):
    v, f = icosphere(subdiv)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    # polynomial "lobes" (body, two ear-like bumps, a tail bump); no transcendental functions
    ear1 = np.maximum(0.0, (0.5 * x + 0.8 * y + 0.2 * z) - 0.75)
    ear2 = np.maximum(0.0, (-0.5 * x + 0.8 * y + 0.2 * z) - 0.75)
    tail = np.maximum(0.0, (-0.9 * z - 0.3 * y) - 0.8)
    ripple = (x * y * z) * (x * x - y * y) * 4.0 + (x * x * z - y * z * z) * 0.35
    r = 0.30 + 0.05 * ripple + 2.2 * ear1 * ear1 * 4.0 + 2.2 * ear2 * ear2 * 4.0 + 1.5 * tail * tail
    p = v * r[:, None]
    p[:, 1] *= 1.15
    p[:, 1] -= p[:, 1].min()           # rest on the floor ...
    p[:, 1] -= 0.25                    # ... which is y = -1 after *8 *0.5
    return p, f


def write_obj(path, subdiv):
    p, f = standin(subdiv)
    with open(path, "w") as fh:
        fh.write("# stand-in mesh (NOT the Stanford bunny): displaced icosphere, subdivision %d, %d triangles\n" % (subdiv, len(f)))
        for q in p:
            fh.write("v %.6f %.6f %.6f\n" % (q[0], q[1], q[2]))
        for t in f:
            fh.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))
    return len(p), len(f)


if __name__ == "__main__":
    sub = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    out = sys.argv[2] if len(sys.argv) > 2 else "standin_%d.obj" % sub
    nv, nf = write_obj(out, sub)
    print("wrote %s: %d vertices, %d triangles" % (out, nv, nf))
