"""TEST INFRASTRUCTURE ONLY — CPU restatement of the device-side synthetic-data recipe (gps_synth_generate):
MakeSyntheticData.R:40-41,145-149 ('likeRealData': x ~ N(0,1), y = exp(t(chol(K)) %*% d1 + sqrt(noise) * d2)) followed by
CensorData.R:36-48 ('NormalLoopSample').  The reference draws from R's Mersenne-Twister through rnorm / sample, seeded from
the wall clock (runSyntheticExp1.R:64): its streams cannot be reproduced, so the device recipe defines its own
counter-based stream (Philox4x32-10, Salmon et al. 2011, the generator cuRAND and numpy also ship) and this file restates
that stream and the recipe bit for bit so the CUDA kernels can be checked against it.
"""
from __future__ import annotations

import math

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
STREAM_X, STREAM_D1, STREAM_D2, STREAM_CENSOR = 0, 1, 2, 3


def philox4x32_10(ctr, key):
    """One Philox4x32-10 block: ctr = (c0, c1, c2, c3), key = (k0, k1) -> four 32-bit words."""
    c0, c1, c2, c3 = (int(v) & MASK for v in ctr)
    k0, k1 = (int(v) & MASK for v in key)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def _u53(hi, lo):
    """(0, 1) uniform from two words: 53 bits, never 0 or 1."""
    return (((hi >> 5) << 26) + (lo >> 6) + 0.5) * (1.0 / 9007199254740992.0)


def normal_pair(index, stream, dataset, seed):
    """Two N(0,1) draws (Box-Muller) of block `index` of `stream` for `dataset`."""
    w = philox4x32_10((index, stream, dataset, 0), (seed & MASK, (seed >> 32) & MASK))
    u1, u2 = _u53(w[0], w[1]), _u53(w[2], w[3])
    r = math.sqrt(-2.0 * math.log(u1))
    return r * math.cos(2.0 * math.pi * u2), r * math.sin(2.0 * math.pi * u2)


def normals(count, stream, dataset, seed):
    out = np.empty(count + (count & 1))
    for i in range(0, count, 2):
        out[i], out[i + 1] = normal_pair(i // 2, stream, dataset, seed)
    return out[:count]


def censor_loop(y, n_censor, cens_mean, cens_sd, dataset, seed, max_draws=None):
    """CensorData.R:36-48 'NormalLoopSample': pick one of the still-uncensored samples uniformly, draw a censoring time
    |N(mean, sd)|, censor if it is shorter than the sample's time; repeat until n_censor samples are censored."""
    y = np.array(y, dtype=np.float64)
    n = y.size
    events = np.ones(n)
    remaining = list(range(n))
    done, draw = 0, 0
    max_draws = max_draws or 4000 * n
    while done < n_censor and draw < max_draws:
        w = philox4x32_10((draw, STREAM_CENSOR, dataset, 0), (seed & MASK, (seed >> 32) & MASK))
        draw += 1
        pick = remaining[min(int(_u53(w[0], w[1]) * len(remaining)), len(remaining) - 1)]     # sample(indicesToCensor, 1)
        u1, u2 = _u53(w[2], w[3]), _u53(w[0] ^ 0x5BD1E995, w[1] ^ 0x9E3779B9)
        z = math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)
        t = abs(cens_mean + cens_sd * z)                                                        # abs(rnorm(1, mean, sd))
        if y[pick] > t:
            y[pick] = t
            events[pick] = 0.0
            remaining.remove(pick)
            done += 1
    return y, events, done


def generate(seed, dataset, n, d, sn2, sf2, ell, jitter, n_censor, cens_mean, cens_sd):
    """The whole recipe for one data set: X (n x d), y before censoring, y after, events."""
    X = normals(n * d, STREAM_X, dataset, seed).reshape(d, n).T.copy()          # column-major fill, like the device
    d1 = normals(n, STREAM_D1, dataset, seed)
    d2 = normals(n, STREAM_D2, dataset, seed)
    Xc = X - X.mean(axis=0)                                                      # distances are translation invariant
    Z = Xc / ell
    s = np.einsum("ij,ij->i", Z, Z)
    r2 = np.maximum(s[:, None] + s[None, :] - 2.0 * (Z @ Z.T), 0.0)
    np.fill_diagonal(r2, 0.0)
    K = sf2 * np.exp(-0.5 * r2) + jitter * np.eye(n)
    L = np.linalg.cholesky(K)
    y0 = np.exp(L @ d1 + math.sqrt(sn2) * d2)                                    # MakeSyntheticData.R:148-149 (zero mean)
    y, ev, done = censor_loop(y0, n_censor, cens_mean, cens_sd, dataset, seed)
    return X, y0, y, ev, done
