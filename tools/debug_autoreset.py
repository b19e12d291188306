"""Find the first divergence between a same-step and a next-step auto-reset batch (debugging aid)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from parity_util import make_env_like
DT = 1.0 / 333.0
n = 64
modeA = int(sys.argv[1]) if len(sys.argv) > 1 else 0
a = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); a.set_seed(5, 0); a.teleport_spline(np.linspace(0, 0.9, n)); a.set_autoreset(modeA)
c = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); c.set_seed(5, 0); c.teleport_spline(np.linspace(0, 0.9, n)); c.set_autoreset(1)
lay = pdref.Layout()
names = {}
for k, (off, ty) in lay.fields.items(): names[off] = k
act = torch.zeros((n, 2), device="cuda"); act[:, 0] = torch.linspace(-1, 1, n, device="cuda"); act[:, 1] = 1.0
ra = torch.zeros(n, device="cuda"); da = torch.zeros(n, device="cuda", dtype=torch.int32)
rc = torch.zeros(n, device="cuda"); dc = torch.zeros(n, device="cuda", dtype=torch.int32)
alive = np.ones(n, bool)
for t in range(1500):
    a.env_step(act, DT, None, ra, da); c.env_step(act, DT, None, rc, dc)
    sa, sc = a.snapshot(), c.snapshot()
    d = da.cpu().numpy(); d2 = dc.cpu().numpy()
    diff = (sa != sc).any(axis=0) & alive
    if diff.any():
        e = int(np.nonzero(diff)[0][0])
        w = np.nonzero(sa[:, e] != sc[:, e])[0]
        print("step", t, "first differing alive env", e, "words", [(int(x), names.get(int(x), "?")) for x in w[:12]], "n words", len(w))
        print("done a", np.nonzero(d)[0], "done c", np.nonzero(d2)[0], "not alive", np.nonzero(~alive)[0])
        print("vals a", sa[w[:6], e].view(np.float32), "c", sc[w[:6], e].view(np.float32))
        break
    alive &= (d == 0)
else:
    print("no divergence among alive envs in 1500 steps; alive left", alive.sum())
