"""Generate tests/golden/*.npz from the oracle (the reference cannot run here)."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as g
g.build()
import wasm_pathtracer_b200 as W
import test_oracle_kat as T
meshes = {3: W.parse_obj(open(os.path.join(g.ROOT, "assets", "_gen", "standin_3.obj")).read(), True)}
ids, vis, dist, rgb, st = T._golden_case(meshes)
np.savez_compressed(os.path.join(HERE, "bunny3_48x32.npz"), ids=ids, visits=vis, dist_bits=dist.view(np.uint32), rgb_bits=rgb.view(np.uint32),
                    stats=np.array([st["rays"], st["paths"], st["node_visits"]], np.int64))
print("wrote golden vectors", st)
