"""BASELINE configs[4]: random-policy rollout through the vectorised env (BatchedProjectDEnv, obs via DLPack), 1024 envs.
The policy is a torch op on the env's own observation tensor: no host round trip."""
import sys, os, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdref
from projectd_core_b200.env import BatchedProjectDEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = BatchedProjectDEnv(pdref.BASE_PATH, num_envs=n, device=0, seed=1, teleport_mode=2, autoreset_mode=1)
obs = env.reset()
stream = torch.cuda.ExternalStream(env.batch.stream(), device=torch.device("cuda", 0))
W = torch.randn(24, 2, device="cuda") * 0.1
with torch.cuda.stream(stream):
    for phase, steps in (("warm-up", 2000), ("timed", 3000)):
        torch.cuda.synchronize(); t0 = time.perf_counter(); ret = torch.zeros(n, device="cuda")
        for t in range(steps):
            a = torch.tanh(obs @ W + torch.randn(n, 2, device="cuda") * 0.5)      # a (random) linear policy on the DLPack'd observation
            obs, rew, term, trunc, _ = env.step(a)
            ret += rew
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%s: %d envs x %d steps in %.2f s -> %.3g env-steps/s (%.1f us per vector step)" % (phase, n, steps, dt, n * steps / dt, 1e6 * dt / steps))
print(env.episode_stats())
