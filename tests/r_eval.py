"""A small evaluator for the subset of R that the reference's hot-path functions are written in (test infrastructure).

R is not installed in this image, but the reference is R *text*: tests/r_parse.py parses it and this module executes
it -- closures with R's argument matching (exact, partial, positional, defaults evaluated lazily in the callee's frame),
vectors with names / dim, lists, 1-based / negative / logical / name indexing, nested replacement (`x[1:2][m] <- v`,
`names(x) <- v`, `x$a$b <- v`), recycling arithmetic, `if` / `for` / `while` / `repeat`, `lapply` / `sapply` / `do.call`.
What R delegates to compiled code is restated as builtins here, and only that: arithmetic (numpy, IEEE doubles like R's),
`dnbinom(log = TRUE)` (50-digit mpmath, rounded once), `pchisq`, `p.adjust`, `optim(method = "BFGS")` (vmmin, SURVEY.md
App. B), `splines::ns`, `lm`.  tests/golden/make_reference_exec_golden.py runs the reference's own function bodies with it
(in the build container, where /root/reference exists) and commits inputs + outputs as golden vectors for the oracle."""
import math

import numpy as np

from tests import r_parse
from tests.r_parse import Node


class RError(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class _Break(Exception):
    pass


class _Next(Exception):
    pass


class RVec:
    """Atomic vector: v is a 1-D numpy array (float64, int64, bool or object for character); matrices keep dim = (nr, nc)
    with column-major data."""
    __slots__ = ("v", "names", "dim", "dimnames", "attrs")

    def __init__(self, v, names=None, dim=None, dimnames=None, attrs=None):
        self.attrs = attrs          # other attributes (attr(x, "name") <- value); dropped by subsetting and arithmetic
        self.v = np.atleast_1d(np.asarray(v))
        if self.v.dtype.kind in "US":
            self.v = self.v.astype(object)
        elif self.v.dtype.kind in "iu" and self.v.dtype != np.int64:
            self.v = self.v.astype(np.int64)
        elif self.v.dtype.kind == "f" and self.v.dtype != np.float64:
            self.v = self.v.astype(np.float64)
        if self.v.ndim == 2 and dim is None:
            dim = self.v.shape
            self.v = self.v.reshape(-1, order="F")
        self.names, self.dim, self.dimnames = names, (tuple(int(d) for d in dim) if dim is not None else None), dimnames

    def __len__(self):
        return len(self.v)

    def mat(self):
        return self.v.reshape(self.dim, order="F")

    def __repr__(self):
        return f"RVec({self.v!r}, names={self.names}, dim={self.dim})"


class RList:
    __slots__ = ("v", "names", "attrs")

    def __init__(self, v=None, names=None):
        self.v = list(v or [])
        self.names = list(names) if names is not None else None
        self.attrs = {}

    def __len__(self):
        return len(self.v)

    def get(self, name, exact=True):
        if self.names is None:
            return None
        if name in self.names:
            return self.v[self.names.index(name)]
        if not exact:      # `$` matches partially when the prefix is unique
            hits = [k for k, n in enumerate(self.names) if n and n.startswith(name)]
            if len(hits) == 1:
                return self.v[hits[0]]
        return None

    def set(self, name, value):
        out = RList(self.v, self.names if self.names is not None else [""] * len(self.v))
        out.attrs = dict(self.attrs)
        if name in out.names:
            k = out.names.index(name)
            if value is None:
                del out.v[k], out.names[k]
            else:
                out.v[k] = value
        elif value is not None:
            out.v.append(value)
            out.names.append(name)
        return out

    def __repr__(self):
        return f"RList({dict(zip(self.names or range(len(self.v)), self.v))!r})"


class Closure:
    def __init__(self, formals, body, env, name="<anonymous>"):
        self.formals, self.body, self.env, self.name = formals, body, env, name


class Promise:
    __slots__ = ("node", "env", "value", "done")

    def __init__(self, node, env):
        self.node, self.env, self.value, self.done = node, env, None, False


class Env:
    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e, e.vars[name]
            e = e.parent
        return None, None

    def set(self, name, value):
        self.vars[name] = value


MISSING = object()


def scalar(x, what="value"):
    if isinstance(x, RVec) and len(x) >= 1:
        return x.v[0]
    raise RError(f"{what}: scalar expected, got {x!r}")


def truth(x):
    if not isinstance(x, RVec) or len(x) == 0:
        raise RError("argument is of length zero")
    v = x.v[0]
    if isinstance(v, float) and math.isnan(v):
        raise RError("missing value where TRUE/FALSE needed")
    return bool(v)


def num(x):
    return RVec(np.asarray(x, dtype=np.float64))


def as_py(x):
    """R value -> plain Python (for JSON): vectors become lists (scalars stay lists of one), matrices lists of rows."""
    if x is None:
        return None
    if isinstance(x, RVec):
        if x.dim is not None:
            return x.mat().tolist()
        return x.v.tolist()
    if isinstance(x, RList):
        if x.names and all(x.names):
            return {n: as_py(v) for n, v in zip(x.names, x.v)}
        return [as_py(v) for v in x.v]
    raise TypeError(x)


def from_py(x):
    """Plain Python / numpy -> R value: dict -> named list, list of non-numbers -> list, arrays -> vectors / matrices."""
    if x is None or isinstance(x, (RVec, RList, Closure)):
        return x
    if isinstance(x, dict):
        return RList([from_py(v) for v in x.values()], list(x.keys()))
    if isinstance(x, (list, tuple)) and any(isinstance(v, (dict, list, tuple, np.ndarray)) or v is None for v in x):
        return RList([from_py(v) for v in x])
    return RVec(np.asarray(x))


class Interp:
    def __init__(self):
        self.globals = Env()
        self.builtins = {}
        self.messages = []
        self.stack = []
        from tests import r_builtins
        r_builtins.install(self)

    # -- loading -------------------------------------------------------------------------------------------------------
    def load(self, src, only_definitions=True):
        """Evaluates the top-level `name <- function ...` / `name <- compiler::cmpfun(other)` definitions of a source text
        (everything else a package file holds -- setClass, setGeneric, roxygen -- is not part of the hot path)."""
        for e in r_parse.parse(src):
            if e.kind == "binop" and e.a[0] in ("<-", "=") and e.a[1].kind == "name":
                rhs = e.a[2]
                if rhs.kind == "function" or (rhs.kind == "call" and rhs.a[0].kind == "ns" and rhs.a[0].a[1] == "cmpfun") \
                        or not only_definitions:
                    val = self.eval(rhs, self.globals)
                    if isinstance(val, Closure):
                        val.name = e.a[1].a[0]
                    self.globals.set(e.a[1].a[0], val)
            elif e.kind == "call" and e.a[0].kind == "name" and e.a[0].a[0] == "setClass":
                self.eval(e, self.globals)            # records the slots new() fills
            elif not only_definitions:
                self.eval(e, self.globals)

    def call(self, fname, *args, **kwargs):
        fn = self.lookup_function(fname, self.globals)
        vals = [(None, from_py(a)) for a in args] + [(k, from_py(v)) for k, v in kwargs.items()]
        return self.apply(fn, vals, self.globals)

    # -- lookup --------------------------------------------------------------------------------------------------------
    def force(self, v):
        if isinstance(v, Promise):
            if not v.done:
                v.value = self.eval(v.node, v.env)
                v.done = True
            return v.value
        return v

    def lookup(self, name, env):
        e, v = env.lookup(name)
        if e is None:
            if name in self.builtins:
                return self.builtins[name]
            raise RError(f"object '{name}' not found")
        if v is MISSING:
            raise RError(f"argument \"{name}\" is missing, with no default")
        return self.force(v)

    def lookup_function(self, name, env):
        e = env
        while e is not None:
            if name in e.vars:
                v = self.force(e.vars[name]) if e.vars[name] is not MISSING else None
                if isinstance(v, Closure) or callable(v):
                    return v
            e = e.parent
        if name in self.builtins:
            return self.builtins[name]
        raise RError(f"could not find function \"{name}\"")

    # -- evaluation ----------------------------------------------------------------------------------------------------
    def eval(self, n, env):
        k = n.kind
        if k == "num":
            t = n.a[0]
            if t.endswith("L"):
                return RVec(np.array([int(t[:-1], 0)], dtype=np.int64))
            return RVec(np.array([float(int(t, 16)) if t[:2].lower() == "0x" else float(t)]))
        if k == "str":
            return RVec(np.array([n.a[0].encode().decode("unicode_escape") if "\\" in n.a[0] else n.a[0]], dtype=object))
        if k == "name":
            return self.lookup(n.a[0], env)
        if k == "paren":
            return self.eval(n.a[0], env)
        if k == "block":
            out = None
            for e in n.a[0]:
                out = self.eval(e, env)
            return out
        if k == "call":
            return self.eval_call(n, env)
        if k == "binop":
            return self.eval_binop(n, env)
        if k == "unop":
            x = self.eval(n.a[1], env)
            if n.a[0] == "-":
                return self._like(x, -x.v)
            if n.a[0] == "+":
                return x
            if n.a[0] == "!":
                return self._like(x, ~self._logical(x.v))
            raise RError("unary " + n.a[0])
        if k == "if":
            if truth(self.eval(n.a[0], env)):
                return self.eval(n.a[1], env)
            return self.eval(n.a[2], env) if n.a[2] is not None else None
        if k == "for":
            seq = self.eval(n.a[1], env)
            items = seq.v if isinstance(seq, RList) else ([] if seq is None else [RVec(np.array([x], dtype=seq.v.dtype)) for x in seq.v])
            for it in items:
                env.set(n.a[0], it)
                try:
                    self.eval(n.a[2], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k in ("while", "repeat"):
            while k == "repeat" or truth(self.eval(n.a[0], env)):
                try:
                    self.eval(n.a[-1], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k == "break":
            raise _Break()
        if k == "next":
            raise _Next()
        if k == "function":
            return Closure(n.a[0], n.a[1], env)
        if k == "dollar":
            obj = self.eval(n.a[0], env)
            if obj is None:
                return None
            if isinstance(obj, RList):
                return obj.get(n.a[1], exact=False)
            if isinstance(obj, Env):
                return self.force(obj.vars.get(n.a[1]))
            raise RError("$ operator is invalid for atomic vectors")
        if k == "at":
            obj = self.eval(n.a[0], env)
            if not isinstance(obj, RList) or "class" not in obj.attrs:
                raise RError("trying to get slot \"%s\" of an object that is not an S4 object" % n.a[1])
            if obj.names is None or n.a[1] not in obj.names:
                raise RError("no slot of name \"%s\" for this object" % n.a[1])
            return obj.v[obj.names.index(n.a[1])]
        if k == "ns":
            if n.a[1] in self.builtins:
                return self.builtins[n.a[1]]
            raise RError(f"{n.a[0]}::{n.a[1]} is not available")
        if k == "index":
            obj = self.eval(n.a[1], env)
            args = [(a, (self.eval(v, env) if v is not None else MISSING)) for a, v in n.a[2]]
            return self.index(obj, n.a[0], args)
        raise RError("cannot evaluate " + k)

    def eval_call(self, n, env):
        callee = n.a[0]
        if callee.kind == "name":
            name = callee.a[0]
            if name == "return":
                raise _Return(self.eval(n.a[1][0][1], env) if n.a[1] and n.a[1][0][1] is not None else None)
            if name == "missing":
                e, v = env.lookup(n.a[1][0][1].a[0])
                return RVec(np.array([v is MISSING]))
            if name == "quote":
                return n.a[1][0][1]
            if name in ("tryCatch", "try"):
                return self._try_catch(n, env)
            if name in ("suppressWarnings", "suppressMessages", "invisible"):
                return self.eval(n.a[1][0][1], env) if n.a[1] else None
            if name == "system.time":
                self.eval(n.a[1][0][1], env)
                return RVec(np.zeros(3), names=["user.self", "sys.self", "elapsed"])
            if name == "switch":
                return self._switch(n, env)
            if name in ("on.exit", "library", "require", "set.seed"):
                return None
            fn = self.lookup_function(name, env)
        else:
            fn = self.eval(callee, env)
        args = []
        for a, v in n.a[1]:
            if v is None:
                continue
            if v.kind == "name" and v.a[0] == "...":
                e, dots = env.lookup("...")
                args.extend(dots or [])
            elif isinstance(fn, Closure):
                args.append((a, Promise(v, env)))
            else:
                args.append((a, self.eval(v, env)))
        return self.apply(fn, args, env)

    def apply(self, fn, args, env):
        if not isinstance(fn, Closure):
            if not callable(fn):
                raise RError("attempt to apply non-function")
            return fn(self, [(a, self.force(v)) for a, v in args], env)
        frame = Env(fn.env)
        names = [f for f, _ in fn.formals]
        bound, rest = {}, []
        for a, v in args:                      # exact names
            if a is not None and a in names and a not in bound and a != "...":
                bound[a] = v
            else:
                rest.append((a, v))
        rest2 = []
        dots_at = names.index("...") if "..." in names else len(names)
        for a, v in rest:                      # unique partial matches (only formals before `...`)
            if a is not None:
                hits = [f for f in names[:dots_at] if f.startswith(a) and f not in bound]
                if len(hits) == 1:
                    bound[hits[0]] = v
                    continue
            rest2.append((a, v))
        dots = []
        free = [f for f in names if f not in bound and f != "..."]
        if "..." in names:
            free = [f for f in names[:dots_at] if f not in bound]
        for a, v in rest2:                     # positional
            if a is None and free:
                bound[free.pop(0)] = v
            elif "..." in names:
                dots.append((a, v))
            else:
                raise RError(f"unused argument in call of {fn.name}: {a if a else '<positional>'}")
        for f, default in fn.formals:
            if f == "...":
                frame.set("...", dots)
            elif f in bound:
                frame.set(f, bound[f])
            elif default is not None:
                frame.set(f, Promise(default, frame))
            else:
                frame.set(f, MISSING)
        self.stack.append(fn.name)
        try:
            return self.eval(fn.body, frame)
        except _Return as r:
            return r.value
        except RError as e:
            if not getattr(e, "rstack", None):
                e.rstack = list(self.stack)
                e.args = (f"{e.args[0] if e.args else ''} [in {' > '.join(self.stack[-6:])}]",)
            raise
        finally:
            self.stack.pop()

    def _try_catch(self, n, env):
        handlers = {a: v for a, v in n.a[1] if a in ("error", "warning", "finally")}
        exprs = [v for a, v in n.a[1] if a is None or a == "expr"]
        try:
            return self.eval(exprs[0], env) if exprs else None
        except RError as e:
            self.messages.append("caught: " + str(e))
            if "error" in handlers:
                h = self.eval(handlers["error"], env)
                cond = RList([RVec(np.array([str(e)], dtype=object))], ["message"])
                return self.apply(h, [(None, cond)], env)
            if n.a[0].a[0] == "try":
                return RVec(np.array(["Error : " + str(e)], dtype=object))
            raise

    def _switch(self, n, env):
        sel = self.eval(n.a[1][0][1], env)
        alts = n.a[1][1:]
        if sel.v.dtype == object:
            key = sel.v[0]
            for k, (a, v) in enumerate(alts):
                if a == key:
                    while v is None:                 # fall through
                        k += 1
                        a, v = alts[k]
                    return self.eval(v, env)
            for a, v in alts:
                if a is None and v is not None:
                    return self.eval(v, env)
            return None
        k = int(sel.v[0])
        return self.eval(alts[k - 1][1], env) if 1 <= k <= len(alts) else None

    # -- arithmetic ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _logical(v):
        if v.dtype == bool:
            return v
        if v.dtype == object:
            raise RError("invalid argument type")
        return v != 0

    @staticmethod
    def _like(x, v):
        return RVec(v, names=x.names if len(v) == len(x) else None, dim=x.dim if len(v) == len(x) else None,
                    dimnames=x.dimnames if len(v) == len(x) else None)

    @staticmethod
    def _recycle(a, b):
        la, lb = len(a), len(b)
        if la == lb or la == 0 or lb == 0:
            return a, b
        n = max(la, lb)
        return (np.resize(a, n) if la < n else a), (np.resize(b, n) if lb < n else b)

    def eval_binop(self, n, env):
        op = n.a[0]
        if op in ("<-", "=", "<<-"):
            value = self.eval(n.a[2], env)
            if isinstance(value, Closure) and n.a[1].kind == "name" and value.name == "<anonymous>":
                value.name = n.a[1].a[0]
            target_env = env
            if op == "<<-":
                e, _ = (env.parent or env).lookup(_root_name(n.a[1]))
                target_env = e or self.globals
            self.assign(n.a[1], value, target_env, env)
            return value
        if op == "->":
            value = self.eval(n.a[1], env)
            self.assign(n.a[2], value, env, env)
            return value
        if op == "&&":
            return RVec(np.array([truth(self.eval(n.a[1], env)) and truth(self.eval(n.a[2], env))]))
        if op == "||":
            return RVec(np.array([truth(self.eval(n.a[1], env)) or truth(self.eval(n.a[2], env))]))
        if op == "~":
            return Node("formula", n.line, n.a[1], n.a[2], env)
        x, y = self.eval(n.a[1], env), self.eval(n.a[2], env)
        if op.startswith("%"):
            fn = self.builtins.get(op)
            if fn is None:
                raise RError("operator " + op)
            return fn(self, [(None, x), (None, y)], env)
        return self.arith(op, x, y)

    def arith(self, op, x, y):
        if not isinstance(x, RVec) or not isinstance(y, RVec):
            raise RError(f"non-numeric argument to binary operator {op}")
        a, b = self._recycle(x.v, y.v)
        shape = x if len(x) >= len(y) else y
        if x.dim is None and y.dim is not None and len(x) == len(y):
            shape = y
        with np.errstate(all="ignore"):
            if op in ("==", "!=", "<", ">", "<=", ">="):
                if a.dtype == object or b.dtype == object:
                    a, b = a.astype(object), b.astype(object)
                r = {"==": np.equal, "!=": np.not_equal, "<": np.less, ">": np.greater, "<=": np.less_equal,
                     ">=": np.greater_equal}[op](a, b)
                return self._like(shape, np.asarray(r, dtype=bool))
            if op in ("&", "|"):
                la, lb = self._logical(a), self._logical(b)
                return self._like(shape, (la & lb) if op == "&" else (la | lb))
            if a.dtype == object or b.dtype == object:
                raise RError(f"non-numeric argument to binary operator {op}")
            if a.dtype == bool:
                a = a.astype(np.int64)
            if b.dtype == bool:
                b = b.astype(np.int64)
            if op == "+":
                r = a + b
            elif op == "-":
                r = a - b
            elif op == "*":
                r = a * b
            elif op == "/":
                r = a.astype(np.float64) / b.astype(np.float64)
            elif op == "^":
                af, bf = a.astype(np.float64), b.astype(np.float64)
                r = np.array([_rpow(p, q) for p, q in zip(af, bf)], dtype=np.float64)
            elif op == ":":
                lo, hi = float(a[0]), float(b[0])
                cnt = int(math.floor(abs(hi - lo) + 1e-10)) + 1
                r = lo + np.arange(cnt) * (1.0 if hi >= lo else -1.0)
                if lo == int(lo):
                    r = r.astype(np.int64)
                return RVec(r)
            else:
                raise RError("operator " + op)
        return self._like(shape, r)

    # -- indexing ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _positions(idx, n, names):
        """0-based positions selected by an R index vector (positive / negative / logical / character)."""
        if idx is MISSING:
            return np.arange(n)
        v = idx.v
        if v.dtype == bool:
            if len(v) < n:
                v = np.resize(v, n)
            return np.nonzero(v)[0]
        if v.dtype == object:
            if names is None:
                raise RError("subscript out of bounds (no names)")
            out = []
            for s in v:
                if s not in names:
                    raise RError(f"subscript out of bounds: {s!r}")
                out.append(list(names).index(s))
            return np.array(out, dtype=np.int64)
        if v.dtype.kind == "f":
            if np.any(np.isnan(v)):
                raise RError("NA subscripts are not supported here")
            v = np.trunc(v).astype(np.int64)
        if len(v) and np.all(v <= 0):
            drop = set((-v[v < 0] - 1).tolist())
            return np.array([k for k in range(n) if k not in drop], dtype=np.int64)
        if np.any(v < 0):
            raise RError("can't mix positive and negative subscripts")
        return v[v > 0] - 1

    def index(self, obj, kind, args):
        drop, exact = True, True
        pos_args = []
        for a, v in args:
            if a == "drop":
                drop = truth(v)
            elif a == "exact":
                exact = truth(v)
            else:
                pos_args.append(v)
        if obj is None:
            return None
        if kind == "[[":
            idx = pos_args[0]
            if isinstance(obj, RList):
                if idx.v.dtype == object:
                    return obj.get(idx.v[0], exact=True)
                k = int(idx.v[0]) - 1
                if not 0 <= k < len(obj):
                    raise RError("subscript out of bounds")
                return obj.v[k]
            p = self._positions(idx, len(obj), obj.names)
            if len(p) != 1 or p[0] >= len(obj):
                raise RError("subscript out of bounds")
            return RVec(obj.v[p])
        if isinstance(obj, RList) and obj.attrs.get("class") == "data.frame" and len(pos_args) == 2:
            return self._index_data_frame(obj, pos_args[0], pos_args[1])
        if isinstance(obj, RList):
            p = self._positions(pos_args[0], len(obj), obj.names)
            return RList([obj.v[k] if k < len(obj) else None for k in p],
                         [obj.names[k] if k < len(obj) else "" for k in p] if obj.names is not None else None)
        if len(pos_args) == 2:
            if obj.dim is None:
                raise RError("incorrect number of dimensions")
            m = obj.mat()
            rn, cn = (obj.dimnames or (None, None))
            pr = self._positions(pos_args[0], obj.dim[0], rn)
            pc = self._positions(pos_args[1], obj.dim[1], cn)
            if (len(pr) and pr.max() >= obj.dim[0]) or (len(pc) and pc.max() >= obj.dim[1]):
                raise RError("subscript out of bounds")
            sub = m[np.ix_(pr, pc)]
            new_rn = [rn[k] for k in pr] if rn is not None else None
            new_cn = [cn[k] for k in pc] if cn is not None else None
            if drop and (len(pr) == 1 or len(pc) == 1):
                names = new_cn if len(pr) == 1 and len(pc) != 1 else (new_rn if len(pc) == 1 and len(pr) != 1 else None)
                return RVec(sub.reshape(-1, order="F"), names=names)
            return RVec(sub.reshape(-1, order="F"), dim=sub.shape, dimnames=(new_rn, new_cn) if (new_rn or new_cn) else None)
        if len(pos_args) == 0:
            return obj
        p = self._positions(pos_args[0], len(obj), obj.names)
        out = np.empty(len(p), dtype=obj.v.dtype if obj.v.dtype != np.int64 and obj.v.dtype != bool else obj.v.dtype)
        inside = p < len(obj)
        if not np.all(inside):
            if obj.v.dtype.kind != "f":
                raise RError("index beyond the vector (NA of a non-double type is not modelled)")
            out = np.full(len(p), np.nan)
            out[inside] = obj.v[p[inside]]
        else:
            out = obj.v[p]
        return RVec(out, names=[obj.names[k] for k in p] if obj.names is not None and np.all(inside) else None)

    def _index_data_frame(self, df, rows, cols):
        """df[rows, cols]: rows may hold NA (a row of NA, as match() produces for absent keys)."""
        rn = df.attrs.get("row.names")
        n = len(df.v[0]) if df.v else 0
        if rows is MISSING:
            pr = np.arange(n, dtype=np.float64)
        elif rows.v.dtype.kind == "f" and np.any(np.isnan(rows.v)):
            pr = rows.v - 1
        else:
            pr = self._positions(rows, n, rn).astype(np.float64)
        pc = self._positions(cols, len(df), df.names)
        out_cols = []
        for k in pc:
            col = df.v[k]
            if col.v.dtype == object:
                vals = np.array([col.v[int(r)] if r == r else None for r in pr], dtype=object)
            else:
                vals = np.array([col.v[int(r)] if r == r else np.nan for r in pr], dtype=np.float64 if np.any(pr != pr) else col.v.dtype)
            out_cols.append(RVec(vals))
        if cols is not MISSING and len(pc) == 1:
            return out_cols[0]
        out = RList(out_cols, [df.names[k] for k in pc])
        out.attrs = {"class": "data.frame",
                     "row.names": [rn[int(r)] if (rn is not None and r == r) else "NA" for r in pr] if rn is not None else None}
        return out

    def assign(self, target, value, target_env, env):
        k = target.kind
        if k in ("name", "str"):
            target_env.set(target.a[0], value)
            return
        if k == "paren":
            return self.assign(target.a[0], value, target_env, env)
        if k == "dollar":
            obj_node = target.a[0]
            cur = self._current(obj_node, target_env, env)
            if isinstance(cur, Env):          # environments are reference objects: modified in place
                cur.vars[target.a[1]] = value
                return
            if cur is None:
                cur = RList()
            if not isinstance(cur, RList):
                raise RError("$<- on an atomic vector")
            return self.assign(obj_node, cur.set(target.a[1], value), target_env, env)
        if k == "at":
            obj_node = target.a[0]
            cur = self.eval(obj_node, env)
            if not isinstance(cur, RList) or cur.names is None or target.a[1] not in cur.names:
                raise RError("no slot of name \"%s\" for this object" % target.a[1])
            out = RList(cur.v, cur.names)
            out.attrs = dict(cur.attrs)
            out.v[out.names.index(target.a[1])] = value
            return self.assign(obj_node, out, target_env, env)
        if k == "index":
            obj_node = target.a[1]
            cur = self._current(obj_node, target_env, env)
            args = [(a, (self.eval(v, env) if v is not None else MISSING)) for a, v in target.a[2]]
            return self.assign(obj_node, self.replace(cur, target.a[0], args, value), target_env, env)
        if k == "call" and target.a[0].kind == "name":
            fname = target.a[0].a[0] + "<-"
            fn = self.lookup_function(fname, env)
            obj_node = target.a[1][0][1]
            cur = self._current(obj_node, target_env, env)
            extra = [(a, self.eval(v, env)) for a, v in target.a[1][1:]]
            return self.assign(obj_node, self.apply(fn, [(None, cur)] + extra + [("value", value)], env), target_env, env)
        raise RError("invalid assignment target")

    def _current(self, node, target_env, env):
        try:
            return self.eval(node, env)
        except RError:
            if node.kind == "name":
                return None
            raise

    def replace(self, cur, kind, args, value):
        pos_args = [v for a, v in args if a not in ("drop", "exact")]
        if kind == "[[" or isinstance(cur, RList):
            if cur is None:
                cur = RList()
            idx = pos_args[0]
            if isinstance(cur, RList):
                if kind == "[" and isinstance(value, RList):
                    p = self._positions(idx, len(cur), cur.names)
                    out = RList(cur.v, cur.names)
                    for j, k2 in enumerate(p):
                        out.v[k2] = value.v[j % len(value)]
                    return out
                if idx.v.dtype == object:
                    return cur.set(idx.v[0], value)
                k = int(idx.v[0]) - 1
                out = RList(cur.v, cur.names)
                while len(out.v) <= k:
                    out.v.append(None)
                    if out.names is not None:
                        out.names.append("")
                if value is None and kind == "[[":
                    del out.v[k]
                    if out.names is not None:
                        del out.names[k]
                else:
                    out.v[k] = value
                return out
        if cur is None:
            cur = RVec(np.array([], dtype=value.v.dtype))
        if not isinstance(value, RVec):
            raise RError("replacement of vector elements by a non-vector")
        data, val = cur.v, value.v
        if data.dtype != val.dtype:            # R's coercion hierarchy: logical < integer < double < character
            order = {"b": 0, "i": 1, "f": 2, "O": 3}
            if order[val.dtype.kind] > order[data.dtype.kind]:
                data = data.astype(val.dtype)
            else:
                val = val.astype(data.dtype)
        data = data.copy()
        if len(pos_args) == 2:
            m = data.reshape(cur.dim, order="F")
            rn, cn = (cur.dimnames or (None, None))
            pr = self._positions(pos_args[0], cur.dim[0], rn)
            pc = self._positions(pos_args[1], cur.dim[1], cn)
            cnt = len(pr) * len(pc)
            if cnt and (len(val) == 0 or cnt % len(val)):
                raise RError("number of items to replace is not a multiple of replacement length")
            if cnt:
                m[np.ix_(pr, pc)] = np.resize(val, cnt).reshape((len(pr), len(pc)), order="F")
            return RVec(m.reshape(-1, order="F"), names=cur.names, dim=cur.dim, dimnames=cur.dimnames)
        p = self._positions(pos_args[0], len(data), cur.names) if pos_args else np.arange(len(data))
        if len(p) == 0:
            return RVec(data, names=cur.names, dim=cur.dim, dimnames=cur.dimnames)
        if len(val) == 0:
            raise RError("replacement has length zero")
        if p.max() >= len(data):
            if data.dtype.kind != "f":
                raise RError("extension of a non-double vector is not modelled")
            data = np.concatenate([data, np.full(p.max() + 1 - len(data), np.nan)])
        data[p] = np.resize(val, len(p))
        return RVec(data, names=cur.names if len(data) == len(cur) else None, dim=cur.dim if len(data) == len(cur) else None,
                    dimnames=cur.dimnames if len(data) == len(cur) else None)


def _rpow(x, y):
    """arithmetic.c R_POW: x^2 is x * x, everything else the C library's pow() (numpy's vector pow is a different routine)."""
    if y == 2.0:
        return x * x
    if x == 1.0 or y == 0.0:
        return 1.0
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf
    except ValueError:
        return math.nan


def _root_name(node):
    while node.kind not in ("name", "str"):
        node = node.a[1] if node.kind == "index" else (node.a[1][0][1] if node.kind == "call" else node.a[0])
    return node.a[0]
