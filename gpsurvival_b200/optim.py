"""Host-side optimisers used where the reference calls ``stats::optim`` (GPRegression.R:28-31).

In an R deployment ``optim`` itself keeps running in R and calls the drop-in
``ForOptimisationLog`` / ``OptimisationGradientLog``; this module exists so that the Python
mirror of the reference interface (gpsurvival_b200/gp.py) can run whole fits without R.
Both routines follow the algorithms R's optim implements (Nash 1990, "Compact Numerical
Methods for Computers": Algorithm 19 Nelder-Mead = R's nmmin, Algorithm 21 variable metric
= R's vmmin) with R's defaults: alpha=1, beta=0.5, gamma=2, abstol=-Inf,
reltol=sqrt(.Machine$double.eps); ``maxit`` counts function evaluations for Nelder-Mead
and iterations for BFGS, as in R.

Pure host logic, numpy only.
"""
from __future__ import annotations

import math

import numpy as np

BIG = 1.0e35
RELTOL = math.sqrt(np.finfo(np.float64).eps)


def _fin(f):
    f = float(f)
    return f if math.isfinite(f) else BIG


# The optimisers are written as GENERATORS that yield the point they want evaluated and receive
# (value, gradient-or-None) back.  A sequential driver (nelder_mead / bfgs below) feeds them from
# fn / gr; the batched driver (gpsurvival_b200/batchfit.py) advances many of them in lock-step
# and serves all their requests with ONE batched GPU evaluation per tick — each fit still follows
# exactly the trajectory the sequential optimiser would.
def nelder_mead_steps(par, maxit=500, abstol=-math.inf, reltol=RELTOL, alpha=1.0, beta=0.5, gamma=2.0):
    """Nelder-Mead as R's nmmin runs it.  The simplex is an (n+1) x n array of vertices with a
    parallel array of function values; ``lo``/``hi`` are the best / worst vertex indices.
    Yields ("f", x); expects (f, _) via send(); returns the optim() result dict."""
    x0 = np.array(par, dtype=np.float64).ravel()
    n = x0.size
    if maxit <= 0:
        f0, _ = yield ("f", x0.copy())
        return {"par": x0, "value": float(f0), "counts": 0, "convergence": 0}
    f0, _ = yield ("f", x0.copy())
    f0 = float(f0)
    if not math.isfinite(f0):
        raise RuntimeError("function cannot be evaluated at initial parameters")
    evals = 1
    convtol = reltol * (abs(f0) + reltol)
    step = max(0.1 * float(np.max(np.abs(x0))) if n else 0.0, 0.0)
    if step == 0.0:
        step = 0.1
    V = np.tile(x0, (n + 1, 1))          # vertices
    F = np.full(n + 1, np.nan)
    F[0] = f0
    size = 0.0
    for j in range(1, n + 1):
        trystep = step
        while V[j, j - 1] == x0[j - 1]:
            V[j, j - 1] = x0[j - 1] + trystep
            trystep *= 10.0
        size += trystep
    oldsize = size
    lo = 0
    need_vertices = True
    fail = 0
    while True:
        if need_vertices:
            for j in range(n + 1):
                if j != lo:
                    fv, _ = yield ("f", V[j].copy())
                    F[j] = _fin(fv)
                    evals += 1
            need_vertices = False
        # best / worst exactly as nmmin scans them (first strict improvement wins)
        vl = F[lo]
        vh = vl
        hi = lo
        for j in range(n + 1):
            if j != lo:
                if F[j] < vl:
                    lo, vl = j, F[j]
                if F[j] > vh:
                    hi, vh = j, F[j]
        if vh <= vl + convtol or vl <= abstol:
            break
        # centroid of all vertices but the worst, accumulated in nmmin's order
        cen = -V[hi].copy()
        for j in range(n + 1):
            cen = cen + V[j]
        cen = cen / n
        xr = (1.0 + alpha) * cen - alpha * V[hi]
        fv, _ = yield ("f", xr.copy())
        fr = _fin(fv)
        evals += 1
        if fr < vl:
            xe = gamma * xr + (1.0 - gamma) * cen
            fv, _ = yield ("f", xe.copy())
            fe = _fin(fv)
            evals += 1
            if fe < fr:
                V[hi], F[hi] = xe, fe
            else:
                V[hi], F[hi] = xr, fr
        else:
            if fr < vh:
                V[hi], F[hi] = xr, fr
            xc = (1.0 - beta) * V[hi] + beta * cen
            fv, _ = yield ("f", xc.copy())
            fc = _fin(fv)
            evals += 1
            if fc < F[hi]:
                V[hi], F[hi] = xc, fc
            elif fr >= vh:
                need_vertices = True
                size = 0.0
                for j in range(n + 1):
                    if j != lo:
                        V[j] = beta * (V[j] - V[lo]) + V[lo]
                        size += float(np.sum(np.abs(V[j] - V[lo])))
                if size < oldsize:
                    oldsize = size
                else:
                    fail = 10
                    break
        if evals > maxit:
            break
    if evals > maxit:
        fail = 1
    return {"par": V[lo].copy(), "value": float(F[lo]), "counts": evals, "convergence": fail}


def bfgs_steps(par, maxit=100, abstol=-math.inf, reltol=RELTOL):
    """Variable-metric (BFGS) minimiser as R's vmmin runs it, with the inverse-Hessian
    approximation kept as a full symmetric matrix.  Yields ("f", x) for line-search values and
    ("g", x) for gradients; expects (f, g) via send() — when the driver already returned a
    gradient with the accepted line-search value it is reused and no ("g", x) request is made."""
    stepredn, acctol, reltest = 0.2, 1.0e-4, 10.0
    b = np.array(par, dtype=np.float64).ravel()
    n = b.size
    if maxit <= 0:
        f, _ = yield ("f", b.copy())
        return {"par": b, "value": float(f), "counts": (0, 0), "convergence": 0}
    f, g_with_f = yield ("f", b.copy())
    f = float(f)
    if not math.isfinite(f):
        raise RuntimeError("initial value in 'vmmin' is not finite")
    fmin = f
    nf = ng = 1
    if g_with_f is None:
        _, g_with_f = yield ("g", b.copy())
    g = np.array(g_with_f, dtype=np.float64).ravel()
    it = 1
    ilast = ng
    Bm = np.eye(n)
    while True:
        if ilast == ng:
            Bm = np.eye(n)
        x_old = b.copy()
        g_old = g.copy()
        t = -(Bm @ g)            # numpy BLAS: same maths as vmmin's loops, O(n^2) without Python overhead
        gradproj = float(t @ g)
        if gradproj < 0.0:
            steplength = 1.0
            accpoint = False
            while True:
                b = x_old + steplength * t
                count = int(np.sum((reltest + x_old) == (reltest + b)))
                if count < n:
                    f, g_with_f = yield ("f", b.copy())
                    f = float(f)
                    nf += 1
                    accpoint = math.isfinite(f) and f <= fmin + gradproj * steplength * acctol
                    if not accpoint:
                        steplength *= stepredn
                if count == n or accpoint:
                    break
            enough = (f > abstol) and abs(f - fmin) > reltol * (abs(fmin) + reltol)
            if not enough:
                count = n
                fmin = f
            if count < n:
                fmin = f
                if g_with_f is None:
                    _, g_with_f = yield ("g", b.copy())
                g = np.array(g_with_f, dtype=np.float64).ravel()
                ng += 1
                it += 1
                t = steplength * t
                c = g - g_old
                d1 = float(t @ c)
                if d1 > 0:
                    xv = Bm @ c
                    d2 = 1.0 + float(xv @ c) / d1
                    Bm = Bm + (d2 * np.outer(t, t) - np.outer(xv, t) - np.outer(t, xv)) / d1
                else:
                    ilast = ng
            else:
                if ilast < ng:
                    count = 0
                    ilast = ng
        else:
            count = 0
            if ilast == ng:
                count = n
            else:
                ilast = ng
        if it >= maxit:
            break
        if ng - ilast > 2 * n:
            ilast = ng
        if count == n and ilast == ng:
            break
    return {"par": b, "value": float(fmin), "counts": (nf, ng), "convergence": 0 if it < maxit else 1}


def lbfgs_steps(par, maxit=100, m=10, reltol=RELTOL):
    """Limited-memory BFGS with vmmin's line search and stopping rules ("L-BFGS using the new
    gradient").  Not one of R's optim methods: it is vmmin (Nash Alg. 21: backtracking by 0.2 from step 1, Armijo
    constant 1e-4, `reltest` no-move test, relative-reduction test, restart from steepest descent before giving up) with
    the dense inverse-Hessian replaced by the two-loop recursion over the last `m` (s, y) pairs, H0 = (s'y / y'y) I.
    No len x len state: what the batched C-side driver keeps on the device is m x len per fit.  This generator is the
    host twin of csrc/lbfgs.cuh (same state machine, same order of operations); protocol as bfgs_steps."""
    stepredn, acctol, reltest = 0.2, 1.0e-4, 10.0
    x = np.array(par, dtype=np.float64).ravel()
    n = x.size
    if maxit <= 0:
        f, _ = yield ("f", x.copy())
        return {"par": x, "value": float(f), "counts": (0, 0), "convergence": 0}
    f, g_with_f = yield ("f", x.copy())
    f = float(f)
    if not math.isfinite(f):
        raise RuntimeError("initial value in 'lbfgs' is not finite")
    if g_with_f is None:
        _, g_with_f = yield ("g", x.copy())
    g = np.array(g_with_f, dtype=np.float64).ravel()
    nf = ng = it = 1
    S, Y, RHO = [], [], []            # oldest first
    gamma = 1.0
    conv = 0
    while True:
        # two-loop recursion: t = -H g
        q = g.copy()
        al = [0.0] * len(S)
        for i in range(len(S) - 1, -1, -1):
            al[i] = RHO[i] * float(S[i] @ q)
            q -= al[i] * Y[i]
        r = (gamma if S else 1.0) * q
        for i in range(len(S)):
            be = RHO[i] * float(Y[i] @ r)
            r += (al[i] - be) * S[i]
        t = -r
        gradproj = float(t @ g)
        if not gradproj < 0.0:
            if S:                     # not a descent direction: forget the history, steepest descent
                S, Y, RHO = [], [], []
                t = -g
                gradproj = float(t @ g)
            if not gradproj < 0.0:
                break
        step = 1.0
        moved = False
        while True:
            b = x + step * t
            if int(np.sum((reltest + x) == (reltest + b))) == n:
                break
            fb, g_with_f = yield ("f", b.copy())
            fb = float(fb)
            nf += 1
            if math.isfinite(fb) and fb <= f + gradproj * step * acctol:
                moved = True
                break
            step *= stepredn
        if not moved:
            if S:                     # restart once from steepest descent (vmmin: reset B = I)
                S, Y, RHO = [], [], []
                continue
            break
        if g_with_f is None:
            _, g_with_f = yield ("g", b.copy())
        gb = np.array(g_with_f, dtype=np.float64).ravel()
        ng += 1
        it += 1
        s, yv = b - x, gb - g
        enough = abs(fb - f) > reltol * (abs(f) + reltol)
        x, g, f = b, gb, fb
        had = len(S) > 0
        sy, yy = float(s @ yv), float(yv @ yv)
        if sy > 0.0:
            S.append(s); Y.append(yv); RHO.append(1.0 / sy)
            gamma = sy / yy
            if len(S) > m:
                S.pop(0); Y.pop(0); RHO.pop(0)
        else:
            S, Y, RHO = [], [], []
        if it >= maxit:
            conv = 1
            break
        if not enough:
            if had:
                S, Y, RHO = [], [], []
            else:
                break
    return {"par": x, "value": float(f), "counts": (nf, ng), "convergence": conv}


def _drive(gen, fn, gr):
    """Run an optimiser generator to completion with plain callables."""
    try:
        kind, x = next(gen)
        while True:
            if kind == "f":
                kind, x = gen.send((fn(x), None))
            else:
                kind, x = gen.send((None, gr(x)))
    except StopIteration as stop:
        return stop.value


def nelder_mead(fn, par, maxit=500, **kw):
    return _drive(nelder_mead_steps(par, maxit=maxit, **kw), fn, None)


def bfgs(fn, gr, par, maxit=100, **kw):
    return _drive(bfgs_steps(par, maxit=maxit, **kw), fn, gr)


def lbfgs(fn, gr, par, maxit=100, **kw):
    return _drive(lbfgs_steps(par, maxit=maxit, **kw), fn, gr)


def optim_steps(par, method="Nelder-Mead", maxit=None):
    """Generator form of optim() for the batched driver."""
    if method == "Nelder-Mead":
        return nelder_mead_steps(par, maxit=500 if maxit is None else maxit)
    if method == "BFGS":
        return bfgs_steps(par, maxit=100 if maxit is None else maxit)
    if method == "L-BFGS":
        return lbfgs_steps(par, maxit=100 if maxit is None else maxit)
    raise NotImplementedError(f"optim method {method!r} (Nelder-Mead and BFGS are restated; L-BFGS is ours)")


def optim(par, fn, gr=None, method="Nelder-Mead", maxit=None):
    """``stats::optim(par, fn, gr, method=, control=list(maxit=))`` for the two methods the
    hot path needs."""
    if method == "Nelder-Mead":
        return nelder_mead(fn, par, maxit=500 if maxit is None else maxit)
    if method == "BFGS":
        if gr is None:
            raise ValueError("BFGS needs a gradient")
        return bfgs(fn, gr, par, maxit=100 if maxit is None else maxit)
    if method == "L-BFGS":
        if gr is None:
            raise ValueError("L-BFGS needs a gradient")
        return lbfgs(fn, gr, par, maxit=100 if maxit is None else maxit)
    raise NotImplementedError(f"optim method {method!r} (Nelder-Mead and BFGS are restated; L-BFGS is ours)")
