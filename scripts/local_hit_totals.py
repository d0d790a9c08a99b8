import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from nprsph_b200.dist import SlabGroup
from oracle import oracle as O
side, world = 256, 4
p = O.dam_break_params(side * world, side, side)
g = SlabGroup.local(world, cell_subdiv=2)
g.apply_params(p)
g.scene_block(side * world, side, side, 0.005, None, 1e-4 * 0.005, 1234)
g.set_paused(False)
g.step(3); g.sync()
for w in range(world):
    s = g.sims[w]; info = g.info(w)
    cap = int(info.cap_own + 2 * info.cap_ghost)
    out = np.empty(cap, np.uint32)
    rc = s.lib.nprsph_debug_read(s._h, 6, out.ctypes.data, out.nbytes)
    assert rc == 0, (rc, s.lib.nprsph_last_error(s._h))
    off = int(info.cap_ghost); off += off & 1
    t = out[off:off + int(info.num_own)].astype(np.int64)
    wt = t[: len(t) // 64 * 64].reshape(-1, 64)
    print(w, "own", len(t), "mean", t.mean().round(2), "min", t.min(), "max", t.max(), "p1/p99", np.percentile(t, [1, 99]),
          "warp max/mean", (wt.max(axis=1) / np.maximum(wt.mean(axis=1), 1)).mean().round(3),
          "pairs equal frac", (t[0::2][:len(t)//2] == t[1::2][:len(t)//2]).mean().round(4))
