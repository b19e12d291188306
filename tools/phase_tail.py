"""Phase-by-phase SM cycles of the quad tick kernel, median warp vs slowest warps (profiling build:
make -C projectd_core_b200 OUT=libpd_b200_dbg.so EXTRA=-DPD_PHASE_CLOCKS).  python tools/phase_tail.py ENVS PREROLL"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PD_DEBUG_CLOCKS"] = "1"
os.environ["PD_B200_LIB"] = os.path.join(ROOT, "projectd_core_b200", "libpd_b200_dbg.so")
import numpy as np, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from projectd_core_b200.env import configure_like_env as make_env_like      # the bench configuration: env that terminates on hit -> detection only, collision warp inside the tick kernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda", 0)
b = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); b.set_seed(1234, 0); b.teleport_mode(2); b.set_autoreset(1)
gen = torch.Generator(device=dev); gen.manual_seed(5)
rew = torch.zeros(n, device=dev); done = torch.zeros(n, device=dev, dtype=torch.int32)
names = ["load+prologue", "suspension", "ray cast", "tyre forces", "thermal", "exchange+aero+steer", "assists+drivetrain", "arb+force sums",
         "build rows", "factor", "schur+solve6", "backsolve+integrate+store", "probes", "nearest point", "(brute)", "probe exchange+spline+locator", "lookahead+scoring"]
for t in range(pre + 4):
    if t % 33 == 0:
        a = (torch.rand((n, 2), device=dev, generator=gen) * 2 - 1).contiguous(); torch.cuda.synchronize()
    b.env_step(a, 1.0 / 333.0, None, rew, done)
    if t in (101, pre, pre + 1, pre + 3):
        c = b.debug_warp_clocks()
        nw = int((c[:4096] > 0).sum())
        tot = c[:nw].astype(np.float64)
        st = c[4096:4096 + nw * 32].reshape(nw, 32)[:, :18].astype(np.float64)
        d = np.diff(st, axis=1)                     # [nw, 17]
        order = np.argsort(tot)
        med = order[nw // 2 - nw // 20: nw // 2 + nw // 20]
        slow = order[-max(1, nw // 100):]
        print("tick %d: %d warps, total p50 %.0f  max %.0f kcycles" % (t, nw, np.median(tot) / 1e3, tot.max() / 1e3))
        print("  %-32s %10s %10s %8s" % ("phase", "median", "slowest1%", "delta"))
        for k, nm in enumerate(names):
            m, s = d[med, k].mean() / 1e3, d[slow, k].mean() / 1e3
            print("  %-32s %10.1f %10.1f %8.1f" % (nm, m, s, s - m))
        print("  %-32s %10.1f %10.1f" % ("sum of phases", d[med].sum(axis=1).mean() / 1e3, d[slow].sum(axis=1).mean() / 1e3))
