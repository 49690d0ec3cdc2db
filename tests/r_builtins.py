"""Builtins of tests/r_eval.py: the part of R that is compiled code in R itself (test infrastructure).  exp / log / pow go
through the C library element by element (R calls the same libm), sum() accumulates in long double like R's rsum();
dnbinom comes from interp.dnbinom (default: 50-digit mpmath rounded once; the golden generator may plug a faster one for
whole fits); optim(method = "BFGS") is vmmin with optim's central-difference gradient."""
import math

import numpy as np

from tests.r_eval import MISSING, Closure, RError, RList, RVec, scalar, truth


def _args(args, names, **defaults):
    """R-style matching of builtin arguments: exact names, then positions."""
    out = dict(defaults)
    pos = [v for a, v in args if a is None]
    for a, v in args:
        if a is not None and a in names:
            out[a] = v
    free = [n for n in names if n not in [a for a, _ in args if a is not None]]
    for n, v in zip(free, pos):
        out[n] = v
    return out


def _f(fn):
    def go(v):
        return np.array([fn(float(x)) for x in v], dtype=np.float64)
    return go


def _safe(fn, dom_err):
    def go(x):
        try:
            return fn(x)
        except (ValueError, OverflowError):
            return dom_err(x)
    return go


_exp = _f(_safe(math.exp, lambda x: math.inf if x > 0 else 0.0))
_log = _f(_safe(math.log, lambda x: -math.inf if x == 0 else math.nan))
_sqrt = _f(_safe(math.sqrt, lambda x: math.nan))


def rsum(v):
    """R's sum() of doubles: a long double running sum (summary.c rsum), rounded to double at the end."""
    if len(v) == 0:
        return 0.0
    if v.dtype.kind != "f":
        return float(np.sum(v.astype(np.int64)))
    return float(np.cumsum(v.astype(np.longdouble))[-1])


def _mpmath_dnbinom_log(x, size, mu):
    import mpmath
    mpmath.mp.dps = 50
    x, size, mu = mpmath.mpf(float(x)), mpmath.mpf(float(size)), mpmath.mpf(float(mu))
    v = mpmath.loggamma(x + size) - mpmath.loggamma(size) - mpmath.loggamma(x + 1) \
        + size * mpmath.log(size / (size + mu)) + x * mpmath.log(mu / (size + mu))
    return float(v)


def vmmin(fn, gr, b, maxit, reltol, abstol=-math.inf):
    """R's vmmin (src/appl/optim.c; Nash, Compact Numerical Methods, Alg. 21), minimising fn.  Returns
    (par, value, fncount, grcount, fail)."""
    stepredn, acctol, reltest = 0.2, 1e-4, 10.0
    n = len(b)
    b = np.array(b, dtype=np.float64)
    if maxit <= 0:
        return b, fn(b), 0, 0, 0
    f = fn(b)
    if not math.isfinite(f):
        raise RError("initial value in 'vmmin' is not finite")
    fmin = f
    funcount = gradcount = 1
    g = gr(b)
    iter_ = 1
    ilast = gradcount
    B = np.zeros((n, n))
    t = np.zeros(n)
    X = np.zeros(n)
    c = np.zeros(n)
    fail = 1
    while True:
        if ilast == gradcount:
            B[:] = 0.0
            for i in range(n):
                B[i, i] = 1.0
        X[:] = b
        c[:] = g
        gradproj = 0.0
        for i in range(n):
            s = 0.0
            for j in range(i + 1):
                s -= B[i, j] * g[j]
            for j in range(i + 1, n):
                s -= B[j, i] * g[j]
            t[i] = s
            gradproj += s * g[i]
        if gradproj < 0.0:
            steplength = 1.0
            accpoint = False
            while True:
                count = 0
                for i in range(n):
                    b[i] = X[i] + steplength * t[i]
                    if reltest + X[i] == reltest + b[i]:
                        count += 1
                if count < n:
                    f = fn(b)
                    funcount += 1
                    accpoint = math.isfinite(f) and (f <= fmin + gradproj * steplength * acctol)
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
                g = gr(b)
                gradcount += 1
                iter_ += 1
                D1 = 0.0
                for i in range(n):
                    t[i] = steplength * t[i]
                    c[i] = g[i] - c[i]
                    D1 += t[i] * c[i]
                if D1 > 0:
                    D2 = 0.0
                    for i in range(n):
                        s = 0.0
                        for j in range(i + 1):
                            s += B[i, j] * c[j]
                        for j in range(i + 1, n):
                            s += B[j, i] * c[j]
                        X[i] = s
                        D2 += s * c[i]
                    D2 = 1.0 + D2 / D1
                    for i in range(n):
                        for j in range(i + 1):
                            B[i, j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1
                else:
                    ilast = gradcount
            else:
                if ilast < gradcount:
                    count = 0
                    ilast = gradcount
        else:
            count = 0
            if ilast == gradcount:
                count = n
            else:
                ilast = gradcount
        if iter_ >= maxit:
            break
        if gradcount - ilast > 2 * n:
            ilast = gradcount
        if count == n and ilast == gradcount:
            fail = 0
            break
    if iter_ < maxit:
        fail = 0
    return b, fmin, funcount, gradcount, fail


def install(I):
    B = I.builtins
    I.dnbinom = _mpmath_dnbinom_log
    I.counters = {"optim_fn": 0}

    def lm_solver(X, y):
        """Least squares by Householder QR (LAPACK through numpy; R's lm.fit uses LINPACK dqrdc2): the same solution, the
        last bits depend on the solver as they would between two builds of R."""
        q, r = np.linalg.qr(X)
        coef = np.linalg.solve(r, q.T @ y)
        return coef, X @ coef
    I.lm_solver = lm_solver

    def reg(*names):
        def deco(fn):
            for nm in names:
                B[nm] = fn
            return fn
        return deco

    def vec_fn(np_fn):
        def fn(I, args, env):
            x = args[0][1]
            return I._like(x, np_fn(x.v.astype(np.float64)))
        return fn

    B["exp"] = vec_fn(_exp)
    B["sqrt"] = vec_fn(_sqrt)
    B["abs"] = lambda I, a, e: I._like(a[0][1], np.abs(a[0][1].v))
    B["floor"] = vec_fn(np.floor)
    B["ceiling"] = vec_fn(np.ceil)
    B["is.finite"] = lambda I, a, e: I._like(a[0][1], np.isfinite(a[0][1].v.astype(np.float64)))
    B["is.infinite"] = lambda I, a, e: I._like(a[0][1], np.isinf(a[0][1].v.astype(np.float64)))

    @reg("log")
    def _(I, args, env):
        a = _args(args, ["x", "base"])
        r = _log(a["x"].v.astype(np.float64))
        if "base" in a:
            r = r / math.log(float(scalar(a["base"])))
        return I._like(a["x"], r)

    B["log2"] = lambda I, a, e: I._like(a[0][1], np.array([math.log2(x) if x > 0 else (-math.inf if x == 0 else math.nan) for x in a[0][1].v], dtype=np.float64))
    B["log10"] = lambda I, a, e: I._like(a[0][1], np.array([math.log10(x) if x > 0 else (-math.inf if x == 0 else math.nan) for x in a[0][1].v], dtype=np.float64))

    @reg("round")
    def _(I, args, env):
        a = _args(args, ["x", "digits"])
        return I._like(a["x"], np.round(a["x"].v.astype(np.float64), int(scalar(a["digits"])) if "digits" in a else 0))

    @reg("c")
    def _(I, args, env):
        if any(isinstance(v, RList) for _, v in args):
            out, names = [], []
            for a, v in args:
                if isinstance(v, RList):
                    out.extend(v.v)
                    names.extend(v.names if v.names is not None else [""] * len(v))
                elif v is not None:
                    out.append(v)
                    names.append(a or "")
            return RList(out, names if any(names) else None)
        parts, names, any_names = [], [], False
        for a, v in args:
            if v is None:
                continue
            parts.append(v.v)
            if a is not None and len(v) == 1:
                names.append(a)
                any_names = True
            elif v.names is not None:
                names.extend(v.names)
                any_names = True
            else:
                names.extend([""] * len(v))
        if not parts:
            return None
        kinds = {p.dtype.kind for p in parts}
        if "O" in kinds:
            parts = [np.array([_as_char(x) for x in p], dtype=object) for p in parts]
        elif "f" in kinds:
            parts = [p.astype(np.float64) for p in parts]
        elif "i" in kinds:
            parts = [p.astype(np.int64) for p in parts]
        return RVec(np.concatenate(parts), names=names if any_names else None)

    @reg("list")
    def _(I, args, env):
        return RList([v for _, v in args], [a or "" for a, _ in args] if any(a for a, _ in args) else None)

    @reg("length")
    def _(I, args, env):
        x = args[0][1]
        return RVec(np.array([0 if x is None else len(x)], dtype=np.int64))

    @reg("rep")
    def _(I, args, env):
        a = _args(args, ["x", "times", "each", "length.out"])
        x = a["x"]
        v = x.v
        if "each" in a:
            v = np.repeat(v, int(scalar(a["each"])))
        if "times" in a:
            t = a["times"].v
            v = np.repeat(v, t.astype(np.int64)) if len(t) == len(v) and len(t) > 1 else np.tile(v, int(t[0]))
        if "length.out" in a:
            v = np.resize(v, int(scalar(a["length.out"])))
        return RVec(v)

    @reg("seq")
    def _(I, args, env):
        a = _args(args, ["from", "to", "by", "length.out"])
        if "from" in a and "to" not in a and "by" not in a and "length.out" not in a:
            n = int(scalar(a["from"])) if len(a["from"]) == 1 else len(a["from"])
            return RVec(np.arange(1, n + 1, dtype=np.int64))
        lo = float(scalar(a["from"])) if "from" in a else 1.0
        if "length.out" in a and "by" not in a:
            n = int(scalar(a["length.out"]))
            hi = float(scalar(a["to"]))
            return RVec(np.array([lo + k * (hi - lo) / (n - 1) for k in range(n)]) if n > 1 else np.array([lo]))
        by = float(scalar(a["by"])) if "by" in a else 1.0
        hi = float(scalar(a["to"]))
        if (hi - lo) * by < 0:
            raise RError("wrong sign in 'by' argument")
        cnt = int(math.floor((hi - lo) / by + 1e-10)) + 1
        r = lo + np.arange(cnt) * by
        if lo == int(lo) and by == int(by) and a.get("from") is not None and a["from"].v.dtype.kind == "i" and ("by" not in a or a["by"].v.dtype.kind == "i"):
            r = r.astype(np.int64)
        return RVec(r)

    B["seq_len"] = lambda I, a, e: RVec(np.arange(1, int(scalar(a[0][1])) + 1, dtype=np.int64))
    B["seq_along"] = lambda I, a, e: RVec(np.arange(1, (0 if a[0][1] is None else len(a[0][1])) + 1, dtype=np.int64))

    def reducer(fn, empty=None):
        def go(I, args, env):
            na_rm = any(a == "na.rm" and truth(v) for a, v in args)
            parts = [v.v for a, v in args if a != "na.rm" and v is not None]
            v = np.concatenate(parts) if parts else np.array([], dtype=np.float64)
            if na_rm and v.dtype.kind == "f":
                v = v[~np.isnan(v)]
            if len(v) == 0:
                if empty is None:
                    raise RError("no non-missing arguments")
                return RVec(np.array([empty]))
            return RVec(np.array([fn(v)]))
        return go

    B["sum"] = reducer(lambda v: rsum(v) if v.dtype.kind == "f" else int(np.sum(v.astype(np.int64))), 0.0)
    B["prod"] = reducer(lambda v: float(np.prod(v.astype(np.float64))), 1.0)
    B["max"] = reducer(lambda v: v.max() if not (v.dtype.kind == "f" and np.any(np.isnan(v))) else np.nan, -math.inf)
    B["min"] = reducer(lambda v: v.min() if not (v.dtype.kind == "f" and np.any(np.isnan(v))) else np.nan, math.inf)
    B["mean"] = reducer(lambda v: _rmean(v.astype(np.float64)), math.nan)
    B["any"] = reducer(lambda v: bool(np.any(v)), False)
    B["all"] = reducer(lambda v: bool(np.all(v)), True)

    def _rmean(v):      # summary.c: long double mean, then one refinement pass
        s = np.cumsum(v.astype(np.longdouble))[-1] / len(v)
        t = np.cumsum((v.astype(np.longdouble) - s))[-1] / len(v)
        return float(s + t)

    B["is.null"] = lambda I, a, e: RVec(np.array([a[0][1] is None]))
    B["is.list"] = lambda I, a, e: RVec(np.array([isinstance(a[0][1], RList)]))
    B["is.numeric"] = lambda I, a, e: RVec(np.array([isinstance(a[0][1], RVec) and a[0][1].v.dtype.kind in "if"]))
    B["is.character"] = lambda I, a, e: RVec(np.array([isinstance(a[0][1], RVec) and a[0][1].v.dtype == object]))
    B["is.function"] = lambda I, a, e: RVec(np.array([isinstance(a[0][1], Closure) or callable(a[0][1])]))

    @reg("is.na")
    def _(I, args, env):
        x = args[0][1]
        if isinstance(x, RList):
            return RVec(np.zeros(len(x), dtype=bool))
        return I._like(x, np.isnan(x.v) if x.v.dtype.kind == "f" else np.zeros(len(x), dtype=bool))

    B["which"] = lambda I, a, e: RVec((np.nonzero(I._logical(a[0][1].v))[0] + 1).astype(np.int64))
    B["which.max"] = lambda I, a, e: RVec(np.array([int(np.nanargmax(a[0][1].v.astype(np.float64))) + 1], dtype=np.int64))
    B["which.min"] = lambda I, a, e: RVec(np.array([int(np.nanargmin(a[0][1].v.astype(np.float64))) + 1], dtype=np.int64))

    def _as_char(x):
        if isinstance(x, (bool, np.bool_)):
            return "TRUE" if x else "FALSE"
        if isinstance(x, (float, np.floating)):
            if x == int(x) and abs(x) < 1e15:
                return str(int(x))
            return repr(float(np.float64(f"{x:.15g}")))
        return str(x)

    B["as.vector"] = lambda I, a, e: a[0][1] if not isinstance(a[0][1], RVec) else RVec(a[0][1].v)
    B["as.double"] = B["as.numeric"] = lambda I, a, e: None if a[0][1] is None else RVec(
        np.array([float(x) for x in a[0][1].v]) if a[0][1].v.dtype == object else a[0][1].v.astype(np.float64), names=None)
    B["as.integer"] = lambda I, a, e: RVec(np.trunc(a[0][1].v.astype(np.float64)).astype(np.int64))
    B["as.logical"] = lambda I, a, e: RVec(I._logical(a[0][1].v))
    B["as.character"] = lambda I, a, e: RVec(np.array([_as_char(x) for x in a[0][1].v], dtype=object))
    B["as.matrix"] = lambda I, a, e: a[0][1] if a[0][1].dim is not None else RVec(a[0][1].v, dim=(len(a[0][1]), 1))
    B["as.list"] = lambda I, a, e: a[0][1] if isinstance(a[0][1], RList) else RList(
        [RVec(np.array([x], dtype=a[0][1].v.dtype)) for x in a[0][1].v], a[0][1].names)
    B["numeric"] = lambda I, a, e: RVec(np.zeros(int(scalar(a[0][1])) if a else 0))
    B["integer"] = lambda I, a, e: RVec(np.zeros(int(scalar(a[0][1])) if a else 0, dtype=np.int64))
    B["logical"] = lambda I, a, e: RVec(np.zeros(int(scalar(a[0][1])) if a else 0, dtype=bool))
    B["character"] = lambda I, a, e: RVec(np.array([""] * (int(scalar(a[0][1])) if a else 0), dtype=object))
    B["identity"] = lambda I, a, e: a[0][1]
    B["cmpfun"] = lambda I, a, e: a[0][1]

    @reg("vector")
    def _(I, args, env):
        a = _args(args, ["mode", "length"])
        n = int(scalar(a["length"])) if "length" in a else 0
        mode = a["mode"].v[0] if "mode" in a else "logical"
        if mode == "list":
            return RList([None] * n)
        return RVec(np.zeros(n, dtype={"numeric": np.float64, "double": np.float64, "integer": np.int64, "logical": bool}.get(mode, object)))

    @reg("names")
    def _(I, args, env):
        x = args[0][1]
        if x is None or x.names is None:
            return None
        return RVec(np.array(list(x.names), dtype=object))

    @reg("names<-")
    def _(I, args, env):
        x, value = args[0][1], args[-1][1]
        nm = None if value is None else [str(s) for s in value.v]
        if isinstance(x, RList):
            out = RList(x.v, nm)
            return out
        return RVec(x.v, names=nm, dim=x.dim, dimnames=x.dimnames)

    @reg("attr")
    def _(I, args, env):
        x, which = args[0][1], args[1][1].v[0]
        if which == "names":
            return B["names"](I, [(None, x)], env)
        if which == "dim":
            return B["dim"](I, [(None, x)], env)
        return (x.attrs or {}).get(which)

    @reg("unlist")
    def _(I, args, env):
        x = args[0][1]
        if x is None or isinstance(x, RVec):
            return x
        flat = []

        def walk(v, prefix):
            if v is None:
                return
            if isinstance(v, RList):
                for k, it in enumerate(v.v):
                    walk(it, (v.names[k] if v.names and v.names[k] else prefix))
            else:
                flat.append((prefix, v))
        walk(x, "")
        if not flat:
            return None
        return B["c"](I, [(None, RVec(v.v, names=([p] * len(v) if p and v.names is None and len(v) == 1 else v.names))) for p, v in flat], env)

    def call_fn(I, fn, vals, env):
        return I.apply(fn, vals, env)

    @reg("lapply", "bplapply")
    def _(I, args, env):
        x, fn = args[0][1], args[1][1]
        extra = [(a, v) for a, v in args[2:] if a != "BPPARAM"]
        items = x.v if isinstance(x, RList) else ([] if x is None else [RVec(np.array([e], dtype=x.v.dtype)) for e in x.v])
        if isinstance(fn, RVec):
            fn = I.lookup_function(fn.v[0], env)
        out = [call_fn(I, fn, [(None, it)] + extra, env) for it in items]
        return RList(out, x.names if x is not None else None)

    @reg("sapply", "vapply")
    def _(I, args, env):
        pos = [(a, v) for a, v in args if a not in ("FUN.VALUE", "simplify", "USE.NAMES")]
        if args and any(a == "FUN.VALUE" for a, _ in args) is False and len(pos) >= 3 and isinstance(pos[2][1], RVec) and False:
            pass
        lst = B["lapply"](I, pos[:2] + [p for p in pos[2:] if p[0] is not None], env)
        vals = lst.v
        if not vals:
            return RList()
        if all(isinstance(v, RVec) and len(v) == 1 for v in vals):
            kinds = {v.v.dtype.kind for v in vals}
            dt = object if "O" in kinds else (np.float64 if "f" in kinds else (np.int64 if "i" in kinds else bool))
            return RVec(np.array([v.v[0] for v in vals], dtype=dt), names=lst.names)
        if all(isinstance(v, RVec) for v in vals) and len({len(v) for v in vals}) == 1 and len(vals[0]) > 1:
            m = np.stack([v.v.astype(np.float64) for v in vals], axis=1)
            return RVec(m.reshape(-1, order="F"), dim=m.shape, dimnames=(vals[0].names, lst.names) if (vals[0].names or lst.names) else None)
        return lst

    @reg("do.call")
    def _(I, args, env):
        fn, lst = args[0][1], args[1][1]
        if isinstance(fn, RVec):
            fn = I.lookup_function(fn.v[0], env)
        vals = [((lst.names[k] or None) if lst.names is not None else None, v) for k, v in enumerate(lst.v)] if lst is not None else []
        return I.apply(fn, vals, env)

    @reg("matrix")
    def _(I, args, env):
        a = _args(args, ["data", "nrow", "ncol", "byrow", "dimnames"])
        data = a["data"].v if "data" in a and a["data"] is not None else np.array([np.nan])
        nr = int(scalar(a["nrow"])) if "nrow" in a else None
        nc = int(scalar(a["ncol"])) if "ncol" in a else None
        if nr is None and nc is None:
            nr, nc = len(data), 1
        elif nr is None:
            nr = -(-len(data) // nc)
        elif nc is None:
            nc = -(-len(data) // nr)
        data = np.resize(data, nr * nc)
        if "byrow" in a and truth(a["byrow"]):
            data = data.reshape((nr, nc)).reshape(-1, order="F")
        return RVec(data, dim=(nr, nc))

    B["dim"] = lambda I, a, e: None if a[0][1] is None or not isinstance(a[0][1], RVec) or a[0][1].dim is None else RVec(np.array(a[0][1].dim, dtype=np.int64))
    def _extent(axis):
        def go(I, args, env):
            x = args[0][1]
            if isinstance(x, RList):
                if x.attrs.get("class") != "data.frame":
                    return None
                return RVec(np.array([len(x.v[0]) if axis == 0 and x.v else (0 if axis == 0 else len(x))], dtype=np.int64))
            return None if x.dim is None else RVec(np.array([x.dim[axis]], dtype=np.int64))
        return go
    B["nrow"], B["ncol"] = _extent(0), _extent(1)
    B["t"] = lambda I, a, e: RVec(a[0][1].mat().T.reshape(-1, order="F"), dim=a[0][1].dim[::-1]) if a[0][1].dim else RVec(a[0][1].v, dim=(1, len(a[0][1])))

    def dimnames_get(axis):
        def go(I, args, env):
            x = args[0][1]
            if isinstance(x, RList):
                if x.attrs.get("class") != "data.frame":
                    return None
                nm = x.attrs.get("row.names") if axis == 0 else x.names
                return None if nm is None else RVec(np.array(list(nm), dtype=object))
            if x is None or x.dimnames is None or x.dimnames[axis] is None:
                return None
            return RVec(np.array(x.dimnames[axis], dtype=object))
        return go

    def dimnames_set(axis):
        def go(I, args, env):
            x, value = args[0][1], args[-1][1]
            if isinstance(x, RList) and x.attrs.get("class") == "data.frame" and value is not None:
                out = RList(x.v, x.names if axis == 0 else [str(s_) for s_ in value.v])
                out.attrs = dict(x.attrs)
                if axis == 0:
                    out.attrs["row.names"] = [str(s_) for s_ in value.v]
                return out
            if value is None and (isinstance(x, RList) or x.dim is None):      # `dimnames<-`: nothing to remove
                return x
            dn = list(x.dimnames or (None, None))
            dn[axis] = None if value is None else [str(s) for s in value.v]
            return RVec(x.v, names=x.names, dim=x.dim, dimnames=tuple(dn))
        return go

    B["rownames"], B["colnames"] = dimnames_get(0), dimnames_get(1)
    B["rownames<-"], B["colnames<-"] = dimnames_set(0), dimnames_set(1)

    def bind(axis):
        def go(I, args, env):
            vals = [v for a, v in args if v is not None and a != "deparse.level"]
            if len(vals) == 1 and isinstance(vals[0], RList):
                vals = vals[0].v
            mats = []
            n = max((v.dim[1 - axis] if v.dim else len(v)) for v in vals)
            for v in vals:
                if v.dim:
                    mats.append(v.mat())
                else:
                    col = np.resize(v.v, n)
                    mats.append(col.reshape((n, 1)) if axis == 1 else col.reshape((1, n)))
            m = np.concatenate(mats, axis=axis)
            return RVec(m.reshape(-1, order="F"), dim=m.shape)
        return go

    B["cbind"], B["rbind"] = bind(1), bind(0)

    def margin(fn, axis):
        def go(I, args, env):
            x = args[0][1]
            na_rm = any(a == "na.rm" and truth(v) for a, v in args)
            m = x.mat().astype(np.float64)
            rows = m if axis == 1 else m.T
            out = []
            for r in rows:
                r = r[~np.isnan(r)] if na_rm else r
                out.append(fn(r))
            names = x.dimnames[0 if axis == 1 else 1] if x.dimnames else None
            return RVec(np.array(out, dtype=np.float64), names=names)
        return go

    B["rowSums"] = margin(lambda r: float(np.cumsum(r.astype(np.longdouble))[-1]) if len(r) else 0.0, 1)
    B["colSums"] = margin(lambda r: float(np.cumsum(r.astype(np.longdouble))[-1]) if len(r) else 0.0, 0)
    B["rowMeans"] = margin(lambda r: float(np.cumsum(r.astype(np.longdouble))[-1] / len(r)) if len(r) else math.nan, 1)
    B["colMeans"] = margin(lambda r: float(np.cumsum(r.astype(np.longdouble))[-1] / len(r)) if len(r) else math.nan, 0)

    @reg("apply")
    def _(I, args, env):
        x, marg, fn = args[0][1], int(scalar(args[1][1])), args[2][1]
        m = x.mat()
        rows = m if marg == 1 else m.T
        vals = [I.apply(fn, [(None, RVec(r.copy()))], env) for r in rows]
        if all(len(v) == 1 for v in vals):
            return RVec(np.array([v.v[0] for v in vals]))
        mm = np.stack([v.v for v in vals], axis=1)
        return RVec(mm.reshape(-1, order="F"), dim=mm.shape)

    @reg("%*%")
    def _(I, args, env):
        x, y = args[0][1], args[1][1]
        a = x.mat().astype(np.float64) if x.dim else None
        b = y.mat().astype(np.float64) if y.dim else None
        if a is None and b is None:
            a, b = x.v.astype(np.float64).reshape(1, -1), y.v.astype(np.float64).reshape(-1, 1)
        elif a is None:
            a = x.v.astype(np.float64).reshape(1, -1) if b.shape[0] == len(x) else x.v.astype(np.float64).reshape(-1, 1)
        elif b is None:
            b = y.v.astype(np.float64).reshape(-1, 1) if a.shape[1] == len(y) else y.v.astype(np.float64).reshape(1, -1)
        if a.shape[1] != b.shape[0]:
            raise RError("non-conformable arguments")
        # R's matprod hands NaN-free doubles to BLAS dgemm / dgemv; the reference BLAS accumulates C[i, j] += B[l, j] * A[i, l]
        # in double, l = 1 .. K in order
        out = np.zeros((a.shape[0], b.shape[1]))
        for l in range(a.shape[1]):
            out = out + np.outer(a[:, l], b[l, :])
        return RVec(out.reshape(-1, order="F"), dim=out.shape)

    @reg("%in%", "is.element")
    def _(I, args, env):
        x, table = args[0][1], args[1][1]
        tb = set(table.v.tolist()) if table is not None else set()
        return RVec(np.array([v in tb for v in x.v.tolist()], dtype=bool))

    @reg("match")
    def _(I, args, env):
        x, table = args[0][1], args[1][1]
        tl = table.v.tolist()
        pos = [tl.index(v) + 1 if v in tl else None for v in x.v.tolist()]
        if any(q is None for q in pos):           # NA_integer_ is modelled as a double NaN
            return RVec(np.array([np.nan if q is None else float(q) for q in pos]))
        return RVec(np.array(pos, dtype=np.int64))

    B["unique"] = lambda I, a, e: RVec(np.array(list(dict.fromkeys(a[0][1].v.tolist())), dtype=a[0][1].v.dtype))
    B["rev"] = lambda I, a, e: RVec(a[0][1].v[::-1].copy())
    B["cumsum"] = lambda I, a, e: RVec(np.cumsum(a[0][1].v))
    B["nchar"] = lambda I, a, e: RVec(np.array([len(s) for s in a[0][1].v], dtype=np.int64))
    B["setdiff"] = lambda I, a, e: RVec(np.array([v for v in dict.fromkeys(a[0][1].v.tolist()) if v not in set(a[1][1].v.tolist())], dtype=a[0][1].v.dtype))

    @reg("sort")
    def _(I, args, env):
        a = _args(args, ["x", "decreasing"])
        v = np.sort(a["x"].v, kind="stable")
        return RVec(v[::-1].copy() if "decreasing" in a and truth(a["decreasing"]) else v)

    @reg("order")
    def _(I, args, env):
        a = _args(args, ["x", "decreasing"])
        v = a["x"].v
        o = np.argsort(-v if "decreasing" in a and truth(a["decreasing"]) else v, kind="stable")
        return RVec((o + 1).astype(np.int64))

    @reg("ifelse")
    def _(I, args, env):
        t, y, n = args[0][1], args[1][1], args[2][1]
        m = I._logical(t.v)
        return I._like(t, np.where(m, np.resize(y.v, len(m)), np.resize(n.v, len(m))))

    @reg("identical")
    def _(I, args, env):
        def same(x, y):
            if x is None or y is None:
                return x is y
            if type(x) is not type(y):
                return False
            if isinstance(x, RVec):
                return x.v.dtype.kind == y.v.dtype.kind and len(x) == len(y) and bool(np.all(x.v == y.v)) and x.dim == y.dim
            if isinstance(x, RList):
                return len(x) == len(y) and x.names == y.names and all(same(p, q) for p, q in zip(x.v, y.v))
            return x is y
        return RVec(np.array([same(args[0][1], args[1][1])]))

    @reg("paste", "paste0")
    def _(I, args, env):
        sep = " "
        collapse = None
        parts = []
        for a, v in args:
            if a == "sep":
                sep = v.v[0]
            elif a == "collapse":
                collapse = None if v is None else v.v[0]
            elif v is not None:
                parts.append([_as_char(x) for x in v.v])
        return parts, sep, collapse

    def paste_impl(default_sep):
        raw = B["paste"]

        def go(I, args, env):
            parts, sep, collapse = raw(I, args, env)
            if not any(a == "sep" for a, _ in args):
                sep = default_sep
            n = max((len(p) for p in parts), default=0)
            out = [sep.join(p[k % len(p)] for p in parts if len(p)) for k in range(n)]
            if collapse is not None:
                out = [collapse.join(out)]
            return RVec(np.array(out, dtype=object))
        return go

    B["paste0"], B["paste"] = paste_impl(""), paste_impl(" ")

    @reg("message", "print", "warning", "cat")
    def _(I, args, env):
        I.messages.append("".join(_as_char(x) for _, v in args if isinstance(v, RVec) for x in v.v))
        return None

    @reg("stop")
    def _(I, args, env):
        raise RError("".join(_as_char(x) for _, v in args if isinstance(v, RVec) for x in v.v))

    B["Sys.time"] = lambda I, a, e: RVec(np.array([0.0]))
    B["exists"] = lambda I, a, e: RVec(np.array([e.lookup(a[0][1].v[0])[0] is not None or a[0][1].v[0] in B]))
    B["TRUE"], B["FALSE"], B["T"], B["F"] = RVec(np.array([True])), RVec(np.array([False])), RVec(np.array([True])), RVec(np.array([False]))
    B["NULL"] = None
    B["NA"], B["NA_real_"], B["NaN"] = RVec(np.array([np.nan])), RVec(np.array([np.nan])), RVec(np.array([np.nan]))
    B["Inf"], B["pi"] = RVec(np.array([np.inf])), RVec(np.array([math.pi]))

    @reg("dnbinom")
    def _(I, args, env):
        a = _args(args, ["x", "size", "prob", "mu", "log"])
        if "mu" not in a or "prob" in a:
            raise RError("dnbinom: only the mu parametrisation is modelled")
        x, size, mu = a["x"].v.astype(np.float64), a["size"].v.astype(np.float64), a["mu"].v.astype(np.float64)
        n = max(len(x), len(size), len(mu)) if min(len(x), len(size), len(mu)) else 0
        x, size, mu = np.resize(x, n), np.resize(size, n), np.resize(mu, n)
        out = np.array([I.dnbinom(x[k], size[k], mu[k]) for k in range(n)], dtype=np.float64)
        if not ("log" in a and truth(a["log"])):
            out = _exp(out)
        return RVec(out)

    @reg("pchisq")
    def _(I, args, env):
        import mpmath
        a = _args(args, ["q", "df", "ncp", "lower.tail", "log.p"])
        lower = truth(a["lower.tail"]) if "lower.tail" in a else True
        mpmath.mp.dps = 50
        out = []
        for q, df in zip(*I._recycle(a["q"].v.astype(np.float64), a["df"].v.astype(np.float64))):
            if math.isnan(q):
                out.append(math.nan)
            elif q <= 0:
                out.append(0.0 if lower else 1.0)
            else:
                out.append(float(mpmath.gammainc(df / 2, 0, q / 2, regularized=True) if lower
                                 else mpmath.gammainc(df / 2, q / 2, mpmath.inf, regularized=True)))
        return RVec(np.array(out))

    @reg("p.adjust")
    def _(I, args, env):
        a = _args(args, ["p", "method", "n"])
        if a["method"].v[0] != "BH":
            raise RError("p.adjust: only BH is modelled")
        p = a["p"].v.astype(np.float64)
        ok = ~np.isnan(p)
        n = int(ok.sum())
        out = np.full(len(p), np.nan)
        ps = p[ok]
        o = np.argsort(-ps, kind="stable")                  # order(p, decreasing = TRUE)
        ro = np.argsort(o, kind="stable")
        adj = np.minimum(1.0, np.minimum.accumulate(n / np.arange(n, 0, -1) * ps[o]))
        out[ok] = adj[ro]
        return RVec(out, names=a["p"].names)

    @reg("optim")
    def _(I, args, env):
        a = _args(args, ["par", "fn", "gr", "method", "lower", "upper", "control", "hessian"])
        if a["method"].v[0] != "BFGS" or a.get("gr") is not None:
            raise RError("optim: only method = 'BFGS' with gr = NULL is modelled")
        known = {"par", "fn", "gr", "method", "lower", "upper", "control", "hessian"}
        extra = [(k, v) for k, v in args if k is not None and k not in known]
        ctl = a.get("control")
        get = (lambda k, d: float(scalar(ctl.get(k))) if ctl is not None and ctl.get(k) is not None else d)
        maxit, reltol, fnscale = int(get("maxit", 100)), get("reltol", math.sqrt(np.finfo(float).eps)), get("fnscale", 1.0)
        par0 = a["par"]
        names = par0.names
        fn = a["fn"]

        def f(b):
            I.counters["optim_fn"] += 1
            val = I.apply(fn, [(None, RVec(np.array(b, dtype=np.float64), names=names))] + extra, env)
            return float(scalar(val)) / fnscale

        def gr(b):
            g = np.empty(len(b))
            for i in range(len(b)):
                eps = 1e-3
                x = np.array(b, dtype=np.float64)
                x[i] = b[i] + eps
                v1 = f(x)
                x[i] = b[i] - eps
                v2 = f(x)
                g[i] = (v1 - v2) / (2 * eps)
                if not math.isfinite(g[i]):
                    raise RError("non-finite finite-difference value [%d]" % (i + 1))
            return g

        b, fmin, fc, gc, fail = vmmin(f, gr, par0.v.astype(np.float64), maxit, reltol)
        return RList([RVec(b, names=names), RVec(np.array([fmin * fnscale])), RVec(np.array([fc, gc], dtype=np.int64), names=["function", "gradient"]),
                      RVec(np.array([fail], dtype=np.int64)), None], ["par", "value", "counts", "convergence", "message"])

    @reg("ns")
    def _(I, args, env):
        from lineagepulse_b200.splines import ns as py_ns
        a = _args(args, ["x", "df", "knots", "intercept", "Boundary.knots"])
        if not ("intercept" in a and truth(a["intercept"])):
            raise RError("ns: only intercept = TRUE is modelled")
        m = py_ns(a["x"].v.astype(np.float64), int(scalar(a["df"])))
        return RVec(m.reshape(-1, order="F"), dim=m.shape)

    @reg("lm")
    def _(I, args, env):
        f = args[0][1]
        if getattr(f, "kind", None) != "formula":
            raise RError("lm: formula expected")
        lhs, rhs, fenv = f.a
        y = I.eval(lhs, fenv).v.astype(np.float64)
        terms = []

        def collect(n):
            if n.kind == "binop" and n.a[0] == "+":
                collect(n.a[1])
                collect(n.a[2])
            else:
                terms.append(n)
        collect(rhs)
        cols, intercept = [], True
        for t in terms:
            if t.kind == "num" and float(t.a[0]) == 0:
                intercept = False
                continue
            v = I.eval(t, fenv)
            cols.append(v.mat().astype(np.float64) if v.dim else v.v.astype(np.float64).reshape(-1, 1))
        X = np.concatenate(([np.ones((len(y), 1))] if intercept else []) + cols, axis=1)
        coef, fitted = I.lm_solver(X, y)
        return RList([RVec(coef), RVec(fitted), RVec(y - fitted)], ["coefficients", "fitted.values", "residuals"])

    @reg("array")
    def _(I, args, env):
        a = _args(args, ["data", "dim"])
        d = a["dim"].v.astype(np.int64)
        if len(d) > 2:
            raise RError("array: more than two dimensions are not modelled")
        n = int(np.prod(d))
        return RVec(np.resize(a["data"].v, n), dim=tuple(d) if len(d) == 2 else None)

    I.classes = {}

    @reg("setClass")
    def _(I, args, env):
        a = _args(args, ["Class", "representation", "slots", "contains"])
        slots = a.get("slots", a.get("representation"))
        I.classes[a["Class"].v[0]] = list(slots.names)
        return None

    @reg("new")
    def _(I, args, env):
        cls = args[0][1].v[0]
        given = {a: v for a, v in args[1:]}
        slots = I.classes.get(cls, list(given))
        for a in given:
            if a not in slots:
                raise RError(f"invalid name for slot of class \u201c{cls}\u201d: {a}")
        out = RList([given.get(sl) for sl in slots], slots)       # unspecified slots: the (NULL) prototype
        out.attrs["class"] = cls
        return out

    @reg("is")
    def _(I, args, env):
        a = _args(args, ["object", "class2"])
        x, cls = a["object"], a["class2"].v[0]
        if isinstance(x, RList):
            return RVec(np.array([x.attrs.get("class") == cls or (cls == "list" and "class" not in x.attrs)]))
        if isinstance(x, RVec):
            kinds = {"character": x.v.dtype == object, "numeric": x.v.dtype.kind in "if", "logical": x.v.dtype == bool,
                     "matrix": x.dim is not None}
            return RVec(np.array([bool(kinds.get(cls, False))]))
        return RVec(np.array([False]))

    B["is.logical"] = lambda I, a, e: RVec(np.array([isinstance(a[0][1], RVec) and a[0][1].v.dtype == bool]))
    B["data.matrix"] = lambda I, a, e: a[0][1]
    B["register"] = B["SerialParam"] = B["MulticoreParam"] = lambda I, a, e: None
    B["packageDescription"] = lambda I, a, e: RVec(np.array(["(executed from source)"], dtype=object))
    B["%%"] = lambda I, a, e: I._like(a[0][1], np.mod(*I._recycle(a[0][1].v, a[1][1].v)))
    B["%/%"] = lambda I, a, e: I._like(a[0][1], np.floor_divide(*I._recycle(a[0][1].v, a[1][1].v)))

    @reg("dimnames")
    def _(I, args, env):
        x = args[0][1]
        if x.dimnames is None:
            return None
        return RList([None if d is None else RVec(np.array(d, dtype=object)) for d in x.dimnames])

    @reg("sparseMatrix")
    def _(I, args, env):
        a = _args(args, ["i", "j", "p", "x", "dims", "dimnames", "symmetric", "index1"])
        one = 1 if ("index1" not in a or truth(a["index1"])) else 0
        nr, nc = (int(v) for v in a["dims"].v)
        m = np.zeros((nr, nc))
        m[a["i"].v.astype(np.int64) - one, a["j"].v.astype(np.int64) - one] = a["x"].v
        dn = a.get("dimnames")
        return RVec(m.reshape(-1, order="F"), dim=(nr, nc),
                    dimnames=None if dn is None else tuple(None if d is None else list(d.v) for d in dn.v))

    @reg("data.frame")
    def _(I, args, env):
        cols = [(a, v) for a, v in args if a not in ("stringsAsFactors", "row.names", "check.names")]
        n = max(len(v) for _, v in cols)
        out = RList([RVec(np.resize(v.v, n)) for _, v in cols], [a for a, _ in cols])
        out.attrs = {"class": "data.frame", "row.names": [str(k + 1) for k in range(n)]}
        return out

    @reg("attr<-")
    def _(I, args, env):
        x, which, value = args[0][1], args[1][1].v[0], args[-1][1]
        if isinstance(x, RList):
            out = RList(x.v, x.names)
            out.attrs = dict(x.attrs)
            out.attrs[which] = value
            return out
        attrs = dict(x.attrs or {})
        attrs[which] = value
        return RVec(x.v, names=x.names, dim=x.dim, dimnames=x.dimnames, attrs=attrs)

    from tests.r_eval import Env
    B["new.env"] = lambda I, a, e: Env(None)
    B["emptyenv"] = lambda I, a, e: None
    B["getOption"] = lambda I, a, e: a[1][1] if len(a) > 1 else None

    @reg(".Call")
    def _(I, args, env):
        if getattr(I, "dot_call", None) is None:
            raise RError(".Call: no native routines are loaded")
        return I.dot_call(args[0][1].v[0], [v for _, v in args[1:]])
    B["invisible"] = lambda I, a, e: a[0][1] if a else None
