"""BASELINE configs[4]: SAC on the batched env (obs via DLPack, 1024 envs), torch-native, no host synchronisation inside the loop.

stable-baselines3 / rl_zoo3 are not in the image, so this is a compact Soft Actor-Critic with the reference's hyper-parameters
(pyprojectd/hyperparams/sac.yml:1-21: MlpPolicy [256, 256], lr 7.3e-4, batch 512, gamma 0.99, tau 0.01, ent_coef auto,
learning_starts 666, observation normalisation) -- enough to time the rollout + update loop the reference's train.py drives and to see
the return move.  Replay buffer, networks, the env's observation / reward / done tensors all live on the GPU; the only host
work per vector step is launching kernels.   python tools/sac_rollout.py [ENVS] [VECTOR_STEPS] [UPDATES_PER_STEP]"""
import os, sys, time
import torch
import torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from projectd_core_b200.assets import default_base
from projectd_core_b200.vector_env import make_vec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
upd = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
env = make_vec(default_base(), num_envs=n, device=0, seed=1, teleport_mode=2)
OBS, ACT, GAMMA, TAU, LR, BATCH, START = 24, 2, 0.99, 0.01, 7.3e-4, 512, 666


def mlp(i, o):
    return nn.Sequential(nn.Linear(i, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, o)).to(dev)


actor = mlp(OBS, 2 * ACT); q1, q2, q1t, q2t = mlp(OBS + ACT, 1), mlp(OBS + ACT, 1), mlp(OBS + ACT, 1), mlp(OBS + ACT, 1)
q1t.load_state_dict(q1.state_dict()); q2t.load_state_dict(q2.state_dict())
log_alpha = torch.zeros((), device=dev, requires_grad=True)
opt_a = torch.optim.Adam(actor.parameters(), lr=LR); opt_q = torch.optim.Adam(list(q1.parameters()) + list(q2.parameters()), lr=LR); opt_al = torch.optim.Adam([log_alpha], lr=LR)
CAP = 1 << 20
buf = dict(o=torch.zeros((CAP, OBS), device=dev), a=torch.zeros((CAP, ACT), device=dev), r=torch.zeros(CAP, device=dev), o2=torch.zeros((CAP, OBS), device=dev), d=torch.zeros(CAP, device=dev))
ptr = 0; size = 0
mean = torch.zeros(OBS, device=dev); var = torch.ones(OBS, device=dev); cnt = 1e-4      # running observation normalisation (norm_obs: True)


def norm(o):
    return torch.clamp((o - mean) / torch.sqrt(var + 1e-8), -10, 10)


def act(o, deterministic=False):
    mu, ls = actor(norm(o)).chunk(2, -1)
    ls = ls.clamp(-20, 2); std = ls.exp()
    u = mu if deterministic else mu + std * torch.randn_like(mu)
    a = torch.tanh(u)
    logp = (-0.5 * ((u - mu) / std) ** 2 - ls - 0.9189385).sum(-1) - torch.log(1 - a * a + 1e-6).sum(-1)
    return a, logp


obs, _ = env.reset()
obs = obs.clone()
ret_sum = torch.zeros((), device=dev); ep_cnt = torch.zeros((), device=dev); run_ret = torch.zeros(n, device=dev)
valid = torch.ones(n, dtype=torch.bool, device=dev)      # False on the step that only delivers a reset observation
torch.cuda.synchronize(); t0 = time.perf_counter(); t_mark = t0
for t in range(steps):
    with torch.no_grad():
        a = (torch.rand((n, ACT), device=dev) * 2 - 1) if t * n < START else act(obs)[0]
    o2, r, term, trunc, _ = env.step(a)
    done = term | trunc
    idx = (ptr + torch.arange(n, device=dev)) % CAP
    w = valid.float()                                     # transitions of reset steps carry zero weight (their action was ignored)
    buf["o"][idx] = obs; buf["a"][idx] = a; buf["r"][idx] = r * w; buf["o2"][idx] = o2; buf["d"][idx] = torch.where(valid, term.float(), torch.ones_like(w))
    ptr = (ptr + n) % CAP; size = min(CAP, size + n)
    with torch.no_grad():                                 # Welford update of the observation statistics
        bm = o2.mean(0); bv = o2.var(0, unbiased=False); tot = cnt + n
        delta = bm - mean; mean += delta * n / tot; var = (var * cnt + bv * n + delta ** 2 * cnt * n / tot) / tot; cnt = tot
    run_ret += r * w
    ret_sum += (run_ret * done).sum(); ep_cnt += done.sum(); run_ret = torch.where(done, torch.zeros_like(run_ret), run_ret)
    valid = ~done
    obs = o2.clone()
    if size >= max(START, BATCH):
        for _ in range(upd):
            j = torch.randint(0, size, (BATCH,), device=dev)
            o, aa, rr, oo2, dd = buf["o"][j], buf["a"][j], buf["r"][j], buf["o2"][j], buf["d"][j]
            alpha = log_alpha.exp().detach()
            with torch.no_grad():
                a2, lp2 = act(oo2)
                x2 = torch.cat([norm(oo2), a2], -1)
                y = rr + GAMMA * (1 - dd) * (torch.min(q1t(x2), q2t(x2)).squeeze(-1) - alpha * lp2)
            x = torch.cat([norm(o), aa], -1)
            lq = ((q1(x).squeeze(-1) - y) ** 2).mean() + ((q2(x).squeeze(-1) - y) ** 2).mean()
            opt_q.zero_grad(set_to_none=True); lq.backward(); opt_q.step()
            an, lp = act(o)
            xn = torch.cat([norm(o), an], -1)
            la = (alpha * lp - torch.min(q1(xn), q2(xn)).squeeze(-1)).mean()
            opt_a.zero_grad(set_to_none=True); la.backward(); opt_a.step()
            lal = -(log_alpha * (lp.detach() + (-ACT))).mean()
            opt_al.zero_grad(set_to_none=True); lal.backward(); opt_al.step()
            with torch.no_grad():
                for p, pt in zip(list(q1.parameters()) + list(q2.parameters()), list(q1t.parameters()) + list(q2t.parameters())):
                    pt.lerp_(p, TAU)
    if (t + 1) % 500 == 0:                                # the only host reads: progress lines
        torch.cuda.synchronize(); now = time.perf_counter()
        print("step %5d: %.3g env-steps/s over the last 500 vector steps; episodes %d, mean return %.2f, alpha %.3f" % (t + 1, 500 * n / (now - t_mark), int(ep_cnt), float(ret_sum / ep_cnt.clamp(min=1)), float(log_alpha.exp())), flush=True)
        t_mark = now; ret_sum.zero_(); ep_cnt.zero_()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("SAC loop: %d envs x %d vector steps (%d gradient updates per step) in %.2f s -> %.3g env-steps/s incl. updates" % (n, steps, upd, dt, n * steps / dt))
print(env.impl.episode_stats())
env.close()
