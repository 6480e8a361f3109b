"""Lock-step batched BFGS: R's ``optim(method='BFGS')`` (vmmin; Nash 1990, Algorithm 21) run for
many independent fits at once, with the dense inverse-Hessian state of ALL fits held in one
[B, n, n] tensor on the GPU.

Why: once an NLML+gradient evaluation costs milliseconds, a host-side dense BFGS becomes the wall
(SURVEY.md §8f rank 1): at C4 (d = 2000 ARD => 2002 hyper-parameters) the O(n^2) update of 128
fits costs ~1.4 s per tick in numpy against ~8 ms of GPU evaluation.  Here the O(n^2) algebra is
three batched tensor ops per tick; the per-fit control flow of vmmin (line-search acceptance,
step reduction, restart rules, termination) is kept exactly, vectorised as boolean masks over the
small per-fit scalars.  torch is plumbing (device tensors, bmm); no oracle import.

``evaluate(P)`` takes a [B, n] float64 numpy array of points and returns (f [B], g [B, n]) —
non-finite f marks a rejected point (e.g. Ky not positive definite).
"""
from __future__ import annotations

import math

import numpy as np
import torch

RELTOL = math.sqrt(np.finfo(np.float64).eps)
STEPREDN, ACCTOL, RELTEST = 0.2, 1.0e-4, 10.0

NEED_DIR, LINESEARCH, DONE = 0, 1, 2


def batch_bfgs(evaluate, x0, maxit=100, abstol=-math.inf, reltol=RELTOL, active=None, device=None):
    """Returns a list of optim()-style dicts (None for inactive fits) and the number of ticks."""
    x0 = np.array(x0, dtype=np.float64)
    B, n = x0.shape
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    act = np.ones(B, dtype=bool) if active is None else np.asarray(active, dtype=bool).copy()
    T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)

    b = T(x0)
    ticks = 0
    if maxit <= 0:
        f, _ = evaluate(x0)
        return [({"par": x0[i].copy(), "value": float(f[i]), "counts": (0, 0), "convergence": 0} if act[i] else None)
                for i in range(B)], 1
    f_np, g_np = evaluate(x0)
    ticks += 1
    f = np.array(f_np, dtype=np.float64)
    if np.any(act & ~np.isfinite(f)):
        raise RuntimeError("initial value in 'vmmin' is not finite")
    g = T(g_np)
    fmin = f.copy()
    nf = np.ones(B, dtype=int)
    ng = np.ones(B, dtype=int)
    it = np.ones(B, dtype=int)
    ilast = ng.copy()
    count = np.zeros(B, dtype=int)
    step = np.ones(B)
    gradproj = np.zeros(B)
    phase = np.where(act, NEED_DIR, DONE)
    Bm = torch.eye(n, dtype=torch.float64, device=dev).repeat(B, 1, 1)
    X = b.clone()
    c = g.clone()
    t = torch.zeros_like(b)
    eye = torch.eye(n, dtype=torch.float64, device=dev)

    def end_of_iteration(idx):
        """vmmin's loop tail for the fits in idx: maxit, periodic restart, termination."""
        for i in idx:
            if it[i] >= maxit:
                phase[i] = DONE
                continue
            if ng[i] - ilast[i] > 2 * n:
                ilast[i] = ng[i]
            phase[i] = DONE if (count[i] == n and ilast[i] == ng[i]) else NEED_DIR

    while np.any(phase != DONE):
        # ---- new search directions (may cascade through the gradproj >= 0 branch without evaluations)
        guard = 0
        while np.any(phase == NEED_DIR):
            idx = np.nonzero(phase == NEED_DIR)[0]
            reset = idx[ilast[idx] == ng[idx]]
            if reset.size:
                Bm[T(reset).long()] = eye
            ti = T(idx).long()
            X[ti] = b[ti]
            c[ti] = g[ti]
            # matvec over the whole batch in place of gathering Bm[ti] (a multi-GB copy at large n)
            t_all = -torch.bmm(Bm, g.unsqueeze(2)).squeeze(2)
            t[ti] = t_all[ti]
            gp_ = (t[ti] * g[ti]).sum(dim=1).cpu().numpy()
            gradproj[idx] = gp_
            down = idx[gp_ < 0.0]
            flat = idx[~(gp_ < 0.0)]
            step[down] = 1.0
            phase[down] = LINESEARCH
            for i in flat:
                count[i] = 0
                if ilast[i] == ng[i]:
                    count[i] = n
                else:
                    ilast[i] = ng[i]
            end_of_iteration(flat)
            guard += 1
            if guard > 4:
                break
        ls = np.nonzero(phase == LINESEARCH)[0]
        if ls.size == 0:
            continue
        # ---- propose b = X + step * t and count the coordinates that did not move
        li = T(ls).long()
        b[li] = X[li] + T(step[ls]).unsqueeze(1) * t[li]
        same = ((RELTEST + X[li]) == (RELTEST + b[li])).sum(dim=1).cpu().numpy()
        count[ls] = same
        need = ls[same < n]
        finished = list(ls[same >= n])            # line search ended without an evaluation
        accepted_with_g = {}
        if need.size:
            f_new, g_new = evaluate(b.cpu().numpy())
            ticks += 1
            f_new = np.array(f_new, dtype=np.float64)
            for i in need:
                f[i] = f_new[i]
                nf[i] += 1
                acc = math.isfinite(f[i]) and f[i] <= fmin[i] + gradproj[i] * step[i] * ACCTOL
                if acc:
                    finished.append(i)
                    accepted_with_g[i] = True
                else:
                    step[i] *= STEPREDN
            g_new_t = T(g_new)
        # ---- fits whose line search is over
        upd = []
        for i in finished:
            enough = (f[i] > abstol) and abs(f[i] - fmin[i]) > reltol * (abs(fmin[i]) + reltol)
            if not enough:
                count[i] = n
                fmin[i] = f[i]
            if count[i] < n:
                fmin[i] = f[i]
                upd.append(i)
            else:
                if ilast[i] < ng[i]:
                    count[i] = 0
                    ilast[i] = ng[i]
        if upd:
            ui_np = np.array(upd)
            ui = T(ui_np).long()
            g[ui] = g_new_t[ui]
            ng[ui_np] += 1
            it[ui_np] += 1
            t[ui] = T(step[ui_np]).unsqueeze(1) * t[ui]
            c[ui] = g[ui] - c[ui]
            D1 = (t[ui] * c[ui]).sum(dim=1)
            D1_np = D1.cpu().numpy()
            pos = D1_np > 0
            if np.any(pos):
                # B += (D2 t t' - X t' - t X') / D1 as ONE in-place rank-3 update of the whole state
                # tensor; fits that do not update this tick get zero coefficients.
                w = torch.zeros(B, dtype=torch.float64, device=dev)
                pi = ui[torch.as_tensor(pos, device=dev)]
                w[pi] = 1.0
                xv_all = torch.bmm(Bm, c.unsqueeze(2)).squeeze(2)
                d1_all = (t * c).sum(dim=1)
                d1_safe = torch.where(w > 0, d1_all, torch.ones_like(d1_all))
                d2_all = 1.0 + (xv_all * c).sum(dim=1) / d1_safe
                s1 = (w * d2_all / d1_safe).unsqueeze(1)
                s2 = (w / d1_safe).unsqueeze(1)
                U = torch.stack([s1 * t, -s2 * xv_all, -s2 * t], dim=2)      # [B, n, 3]
                V = torch.stack([t, t, xv_all], dim=1)                       # [B, 3, n]
                Bm.baddbmm_(U, V)
                X[pi] = xv_all[pi]
            for k, i in enumerate(upd):
                if not pos[k]:
                    ilast[i] = ng[i]
        end_of_iteration(finished)

    b_np = b.cpu().numpy()
    out = []
    for i in range(B):
        if not act[i]:
            out.append(None)
        else:
            out.append({"par": b_np[i].copy(), "value": float(fmin[i]), "counts": (int(nf[i]), int(ng[i])),
                        "convergence": 0 if it[i] < maxit else 1})
    return out, ticks
