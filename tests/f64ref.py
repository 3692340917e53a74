"""float64 restatement of the PPO update phase (numpy) — the ARBITER for fp32 summation-order questions.

TEST INFRASTRUCTURE ONLY (like oracle/).  Same algorithm as oracle/ppo_oracle.c::orc_update
(/root/reference/src/ppo.cu:326-369 GAE, :395-444 schedule, src/loss.cu:5-23, src/policy.cu:67-111,
src/ppo.cu:82-107, src/adam.cu:53-74), every product and sum carried in float64, the hyper-parameters
taken as the float32 constants the reference passes around (so the only difference to the fp32 paths
is rounding inside the arithmetic).  The permutations are an INPUT (the ones the oracle logged), so
all three implementations (reference-order fp32 oracle, GPU, this file) visit identical minibatches.

SURVEY.md §8d: where fp32 summation order decides the last digits (Adam's first steps turn a gradient
that is fp32 noise around zero into +-lr), the float64 result is the arbiter and the reference-order
oracle's own deviation from it is reported next to the GPU's.
"""
import numpy as np

f64 = np.float64


def _split(params, sizes):
    out, o = [], 0
    for i in range(len(sizes) - 1):
        k = sizes[i] * sizes[i + 1]
        W = params[o:o + k].reshape(sizes[i + 1], sizes[i])
        o += k
        b = params[o:o + sizes[i + 1]]
        o += sizes[i + 1]
        out.append((W, b))
    return out


def _act(x, a):
    return np.maximum(x, 0) if a == "relu" else np.tanh(x) if a == "tanh" else x


def _act_grad(y, g, a):
    return g * (y > 0) if a == "relu" else g * (1 - y * y) if a == "tanh" else g


def forward(params, sizes, acts, x):
    hs = [x]
    for (W, b), a in zip(_split(params, sizes), acts):
        hs.append(_act(hs[-1] @ W.T + b, a))
    return hs


def backward(params, sizes, acts, hs, g):
    layers = _split(params, sizes)
    grads = [None] * len(layers)
    for i in range(len(layers) - 1, -1, -1):
        g = _act_grad(hs[i + 1], g, acts[i])
        grads[i] = np.concatenate([(g.T @ hs[i]).ravel(), g.sum(0)])
        g = g @ layers[i][0]
    return np.concatenate(grads)


class Adam64:
    def __init__(self, n):
        self.m, self.v, self.t = np.zeros(n, f64), np.zeros(n, f64), 0

    def step(self, w, g, lr):
        b1, b2 = f64(np.float32(0.9)), f64(np.float32(0.999))
        self.t += 1
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        self.m = b1 * self.m + (1 - b1) * g
        self.v = b2 * self.v + (1 - b2) * g * g
        w -= (lr / bc1) * self.m / (np.sqrt(self.v / bc2) + 1e-8)


def gae(reward, v, v_next, term, trunc, gamma, lam):
    n = reward.shape[0]
    delta = reward + gamma * v_next * (1 - term) - v
    adv = np.zeros(n + 1, f64)
    c = gamma * lam * (1 - np.maximum(term, trunc))
    for i in range(n - 1, -1, -1):
        adv[i] = delta[i] + c[i] * adv[i + 1]
    adv = adv[:n]
    tgt = v + adv
    mean = adv.mean()
    std = np.sqrt(((adv - mean) ** 2).mean())
    return (adv - mean) / (std + 1e-8), tgt


def update(sizes, acts, mu, v, log_std, b, perms, mb, npol, nval, lr_policy=3e-4, lr_v=3e-4, lam=0.95, eps=0.2,
           ent=0.0, gamma=0.99):
    """Returns (mu, v, log_std, advantage, adv_target) after the update, all float64."""
    c = lambda x: f64(np.float32(x))   # noqa: E731
    lr_policy, lr_v, lam, eps, ent, gamma = c(lr_policy), c(lr_v), c(lam), c(eps), c(ent), c(gamma)
    sizes_v = list(sizes[:-1]) + [1]
    A = sizes[-1]
    mu, v, log_std = mu.astype(f64).copy(), v.astype(f64).copy(), log_std.astype(f64).copy()
    st, ns = b["state"].astype(f64), b["next_state"].astype(f64)
    act, lp_old_all = b["action"].astype(f64), b["logprob"].astype(f64)
    term, trunc = b["terminated"].astype(f64), b["truncated"].astype(f64)
    n = st.shape[0]
    adv, tgt = gae(b["reward"].astype(f64), forward(v, sizes_v, acts, st)[-1][:, 0], forward(v, sizes_v, acts, ns)[-1][:, 0],
                   term, trunc, gamma, lam)
    nb = n // mb
    ad_v, ad_mu, ad_ls = Adam64(v.size), Adam64(mu.size), Adam64(A)
    e = 0
    for _ in range(nval):
        perm = perms[e]
        e += 1
        for k in range(nb):
            rows = perm[k * mb:(k + 1) * mb]
            hs = forward(v, sizes_v, acts, st[rows])
            g = 2 * (hs[-1][:, 0] - tgt[rows]) / mb
            ad_v.step(v, backward(v, sizes_v, acts, hs, g[:, None]), lr_v)
    for _ in range(npol):
        perm = perms[e]
        e += 1
        for k in range(nb):
            rows = perm[k * mb:(k + 1) * mb]
            hs = forward(mu, sizes, acts, st[rows])
            m_ = hs[-1]
            z = (act[rows] - m_) / np.exp(log_std)
            lp = -0.5 * A * np.log(2 * np.pi) - (log_std + 0.5 * z * z).sum(1)
            ratio = np.exp(lp - lp_old_all[rows])
            a_ = adv[rows]
            pos = a_ > 0
            keep = np.where(pos, ratio <= 1 + eps, ratio >= 1 - eps)
            g = -(keep * a_ * ratio) / mb
            e2 = np.exp(-2 * log_std)
            diff = act[rows] - m_
            gmu = diff * e2 * g[:, None]
            gls = ((-1 + diff * diff * e2) * g[:, None]).sum(0) - ent
            grads = backward(mu, sizes, acts, hs, gmu)
            ad_ls.step(log_std, gls, lr_policy)
            ad_mu.step(mu, grads, lr_policy)
    return mu, v, log_std, adv, tgt
