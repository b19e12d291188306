"""Driver for ncu / timing experiments: N envs stepped through pd_env_step (actions resampled every 33 ticks, auto-reset),
i.e. the bench workload without the bench's bookkeeping.  python tools/prof_env.py ENVS TICKS [e2e]"""
import sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200 import Batch
from projectd_core_b200.env import configure_like_env as make_env_like      # the bench configuration: env that terminates on hit -> detection only, collision warp inside the tick kernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 400
mode = sys.argv[3] if len(sys.argv) > 3 else "dev"
dev = torch.device("cuda", 0)
b = make_env_like(Batch(pdref.BASE_PATH, n_envs=n, device=0)); b.set_seed(1234, 0); b.teleport_mode(2); b.set_autoreset(int(os.environ.get("PD_AUTORESET", "1")))
gen = torch.Generator(device=dev); gen.manual_seed(5)
rew = torch.zeros(n, device=dev); done = torch.zeros(n, device=dev, dtype=torch.int32)
if mode == "dev":
    for t in range(ticks):
        if t % 33 == 0:
            a = (torch.rand((n, 2), device=dev, generator=gen) * 2 - 1).contiguous(); torch.cuda.synchronize()
        b.env_step(a, 1.0 / 333.0, None, rew, done)
    b.sync()
    print("done", b.launch_count())
else:
    h_act = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    h_obs = torch.empty((n, 24), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_done = torch.empty(n, dtype=torch.int32).pin_memory()
    for rep in range(3):
        T = {"copy": 0.0, "call": 0.0, "read": 0.0}
        t00 = time.perf_counter()
        for t in range(ticks):
            if t % 33 == 0:
                a = (torch.rand((n, 2), generator=None) * 2 - 1).contiguous()
            t0 = time.perf_counter(); h_act.copy_(a); t1 = time.perf_counter()
            b.env_step_host(h_act, 1.0 / 333.0, h_obs, h_rew, h_done); t2 = time.perf_counter()
            x = float(h_rew[0]); t3 = time.perf_counter()
            T["copy"] += t1 - t0; T["call"] += t2 - t1; T["read"] += t3 - t2
        tot = time.perf_counter() - t00
        print("e2e n=%d: %.1f us/step  (copy %.1f call %.1f read %.1f)  -> %.3g car-ticks/s" % (n, 1e6 * tot / ticks, 1e6 * T["copy"] / ticks, 1e6 * T["call"] / ticks, 1e6 * T["read"] / ticks, n * ticks / tot))
