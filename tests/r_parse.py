"""A parser for R source (test infrastructure).  R is not installed in this image, so r_shim/R/lpb200_wrappers.R has never
been parsed by R; this is a restatement of R's grammar (src/main/gram.y of R: operator precedences, newline rules, `if` /
`for` / `while` / `repeat` / `function`, calls, `[` / `[[` indexing, `$` / `@` / `::`) as a Pratt parser that builds a
syntax tree, plus the analyses a never-run file needs: free variables of every function, call sites against the formals
of the callee, element / slot names read against names that are set somewhere.  The parser itself is validated on the
reference's own 6 800 lines of R (every file must parse: tests/test_host_logic.py::test_r_parser_accepts_the_reference)
and on mutated sources it must reject."""
import re


class Node:
    __slots__ = ("kind", "a", "line")

    def __init__(self, kind, line, *a):
        self.kind, self.a, self.line = kind, a, line

    def __repr__(self):
        return f"{self.kind}{self.a!r}"


_OPS = re.compile(r"%[^%\n]*%|<<-|->>|\|>|<-|->|<=|>=|==|!=|&&|\|\||:::|::|[-+*/^<>=!&|~?:$@,;\\]")
_NUM = re.compile(r"0[xX][0-9a-fA-F]+[Li]?|(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?[Li]?")


def lex(src):
    """[(kind, text, line)]: kind in str, name, num, op, open, close, nl; comments and blanks dropped, newlines kept."""
    out, i, n, line = [], 0, len(src), 1
    while i < n:
        c = src[i]
        if c == "\n":
            out.append(("nl", "\n", line))
            line += 1
            i += 1
        elif c in " \t\r\f":
            i += 1
        elif c == "#":
            while i < n and src[i] != "\n":
                i += 1
        elif c in "\"'`":
            j, l0 = i + 1, line
            while j < n and src[j] != c:
                if src[j] == "\\" and c != "`":
                    j += 1
                if j < n and src[j] == "\n":
                    line += 1
                j += 1
            if j >= n:
                raise SyntaxError(f"line {l0}: unterminated {c} string")
            out.append(("name" if c == "`" else "str", src[i + 1:j], l0))
            i = j + 1
        elif c.isalpha() or (c == "." and not (i + 1 < n and src[i + 1].isdigit())):
            j = i + 1
            while j < n and (src[j].isalnum() or src[j] in "._"):
                j += 1
            out.append(("name", src[i:j], line))
            i = j
        elif c.isdigit() or c == ".":
            m = _NUM.match(src, i)
            if not m:
                raise SyntaxError(f"line {line}: malformed number")
            out.append(("num", m.group(0), line))
            i = m.end()
        elif c in "([{":
            if src.startswith("[[", i):
                out.append(("open", "[[", line))
                i += 2
            else:
                out.append(("open", c, line))
                i += 1
        elif c in ")]}":
            out.append(("close", c, line))
            i += 1
        else:
            m = _OPS.match(src, i)
            if not m:
                raise SyntaxError(f"line {line}: unexpected character {c!r}")
            out.append(("op", m.group(0), line))
            i = m.end()
    out.append(("eof", "", line))
    return out


# binary operators: (left binding power, right binding power); gram.y's %left / %right table, low to high
_BIN = {"?": (1, 2), "=": (4, 3), "<-": (6, 5), "<<-": (6, 5), "->": (7, 8), "->>": (7, 8), "~": (9, 10),
        "||": (11, 12), "|": (11, 12), "&&": (13, 14), "&": (13, 14),
        "==": (17, 18), "!=": (17, 18), "<": (17, 18), ">": (17, 18), "<=": (17, 18), ">=": (17, 18),
        "+": (19, 20), "-": (19, 20), "*": (21, 22), "/": (21, 22), "|>": (23, 24), ":": (25, 26), "^": (30, 29)}
_SPECIAL_BP = (23, 24)          # %any%
_UNARY = {"!": 15, "~": 10, "-": 27, "+": 27, "?": 2}
_POSTFIX_BP = 33                # ( [ [[ $ @ bind tighter than everything else
_KEYWORDS = {"if", "else", "for", "while", "repeat", "function", "break", "next", "in"}
_COMPARE = {"==", "!=", "<", ">", "<=", ">="}


class Parser:
    def __init__(self, src):
        self.t = lex(src)
        self.i = 0
        self.ctx = ["top"]      # 'top' / 'brace': a newline ends an expression; 'paren': newlines are blanks

    # -- token helpers ---------------------------------------------------------------------------------------------
    def _skip_nl(self):
        while self.t[self.i][0] == "nl":
            self.i += 1

    def peek(self):
        if self.ctx[-1] == "paren":
            self._skip_nl()
        return self.t[self.i]

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def err(self, msg, tok=None):
        tok = tok or self.t[self.i]
        raise SyntaxError(f"line {tok[2]}: {msg} (at {tok[1]!r})")

    def expect(self, kind, text):
        tok = self.next()
        if tok[0] != kind or tok[1] != text:
            self.err(f"expected {text!r}", tok)
        return tok

    # -- programme / blocks ----------------------------------------------------------------------------------------
    def program(self):
        out = []
        while True:
            self._skip_sep()
            if self.t[self.i][0] == "eof":
                return out
            out.append(self.expr(0))
            tok = self.t[self.i]
            if tok[0] not in ("nl", "eof") and tok[:2] != ("op", ";"):
                self.err("unexpected token after a complete expression", tok)

    def _skip_sep(self):
        while self.t[self.i][0] == "nl" or self.t[self.i][:2] == ("op", ";"):
            self.i += 1

    def block(self, line):
        self.ctx.append("brace")
        out = []
        while True:
            self._skip_sep()
            tok = self.t[self.i]
            if tok[:2] == ("close", "}"):
                self.i += 1
                break
            if tok[0] == "eof":
                self.err("'{' opened on line %d never closed" % line, tok)
            out.append(self.expr(0))
            tok = self.t[self.i]
            if tok[0] != "nl" and tok[:2] not in (("op", ";"), ("close", "}")):
                self.err("unexpected token after a complete expression", tok)
        self.ctx.pop()
        return Node("block", line, out)

    # -- expressions -----------------------------------------------------------------------------------------------
    def expr(self, rbp):
        left = self.prefix()
        while True:
            tok = self.peek()
            kind, text = tok[0], tok[1]
            if kind == "open" and text in ("(", "[", "[["):
                if _POSTFIX_BP < rbp:
                    break
                left = self.postfix(left)
                continue
            if kind != "op":
                break
            if text in ("$", "@"):
                self.next()
                self._skip_nl()
                rhs = self.next()
                if rhs[0] in ("name", "str"):
                    left = Node("dollar" if text == "$" else "at", tok[2], left, rhs[1])
                elif rhs[:2] == ("open", "(") and text == "$":
                    self.i -= 1
                    self.ctx.append("paren")
                    self.next()
                    inner = self.expr(0)
                    self.expect("close", ")")
                    self.ctx.pop()
                    left = Node("dollar", tok[2], left, inner)
                else:
                    self.err("name expected after " + text, rhs)
                continue
            if text in ("::", ":::"):
                self.err("namespace operator after a non-name", tok)
            bp = _SPECIAL_BP if text.startswith("%") else _BIN.get(text)
            if bp is None or bp[0] < rbp:
                break
            self.next()
            self._skip_nl()          # an operator at the end of a line continues the expression
            right = self.expr(bp[1])
            left = Node("binop", tok[2], text, left, right)
            if text in _COMPARE:     # %nonassoc: a second comparison at the same level is a syntax error in R
                nxt = self.peek()
                if nxt[0] == "op" and nxt[1] in _COMPARE and _BIN[nxt[1]][0] >= rbp:
                    self.err("comparison operators do not associate", nxt)
        return left

    def prefix(self):
        tok = self.next()
        kind, text, line = tok
        if kind == "num":
            return Node("num", line, text)
        if kind == "str":
            nxt = self.peek()
            if nxt[0] == "op" and nxt[1] in ("::", ":::"):
                return self._namespace(text, line)
            return Node("str", line, text)
        if kind == "name":
            if text == "function":
                return self.function(line)
            if text == "if":
                return self.if_(line)
            if text == "for":
                self.ctx.append("paren")
                self.expect("open", "(")
                var = self.next()
                if var[0] != "name":
                    self.err("loop variable expected", var)
                self.expect("name", "in")
                seq = self.expr(0)
                self.expect("close", ")")
                self.ctx.pop()
                self._skip_nl()
                return Node("for", line, var[1], seq, self.expr(3))
            if text == "while":
                cond = self.condition()
                self._skip_nl()
                return Node("while", line, cond, self.expr(3))
            if text == "repeat":
                self._skip_nl()
                return Node("repeat", line, self.expr(3))
            if text in ("break", "next"):
                return Node(text, line)
            if text in ("else", "in"):
                self.err("unexpected " + text, tok)
            nxt = self.peek()
            if nxt[0] == "op" and nxt[1] in ("::", ":::"):
                return self._namespace(text, line)
            return Node("name", line, text)
        if kind == "open" and text == "(":
            self.ctx.append("paren")
            inner = self.expr(0)
            self.expect("close", ")")
            self.ctx.pop()
            return Node("paren", line, inner)
        if kind == "open" and text == "{":
            return self.block(line)
        if kind == "op" and text == "\\":
            return self.function(line)
        if kind == "op" and text in _UNARY:
            self._skip_nl()
            return Node("unop", line, text, self.expr(_UNARY[text]))
        self.err("expression expected", tok)

    def _namespace(self, pkg, line):
        op = self.next()
        rhs = self.next()
        if rhs[0] not in ("name", "str"):
            self.err("name expected after " + op[1], rhs)
        return Node("ns", line, pkg, rhs[1])

    def condition(self):
        self.ctx.append("paren")
        self.expect("open", "(")
        cond = self.expr(0)
        self.expect("close", ")")
        self.ctx.pop()
        return cond

    def if_(self, line):
        cond = self.condition()
        self._skip_nl()
        yes = self.expr(3)
        no = None
        save = self.i
        if self.ctx[-1] != "top":        # inside { } or ( ) an else may follow on the next line; at top level it may not
            self._skip_nl()
        if self.t[self.i][:2] == ("name", "else"):
            self.i += 1
            self._skip_nl()
            no = self.expr(3)
        else:
            self.i = save
        return Node("if", line, cond, yes, no)

    def function(self, line):
        self.ctx.append("paren")
        self.expect("open", "(")
        formals = []
        if self.peek()[:2] != ("close", ")"):
            while True:
                name = self.next()
                if name[0] != "name" or name[1] in _KEYWORDS:
                    self.err("formal argument name expected", name)
                default = None
                if self.peek()[:2] == ("op", "="):
                    self.next()
                    default = self.expr(5)
                formals.append((name[1], default))
                tok = self.next()
                if tok[:2] == ("close", ")"):
                    break
                if tok[:2] != ("op", ","):
                    self.err("',' or ')' expected in the formals", tok)
        else:
            self.next()
        self.ctx.pop()
        self._skip_nl()
        return Node("function", line, formals, self.expr(3))

    def postfix(self, left):
        tok = self.next()
        text, line = tok[1], tok[2]
        self.ctx.append("paren")
        args = []
        while True:
            nxt = self.peek()
            if nxt[0] == "close":
                break
            if nxt[:2] == ("op", ","):
                self.next()
                args.append((None, None))            # empty argument: x[, 1]
                if self.peek()[0] == "close":
                    args.append((None, None))
                continue
            name = None
            if nxt[0] in ("name", "str") and self.t[self.i + 1 + self._nl_ahead(self.i + 1)][:2] == ("op", "=") and \
                    nxt[1] not in ("function", "if", "for", "while", "repeat"):
                self.next()
                self.next()
                name = nxt[1]
                if self.peek()[0] == "close" or self.peek()[:2] == ("op", ","):
                    args.append((name, None))        # name = (empty)
                    if self.peek()[:2] == ("op", ","):
                        self.next()
                        if self.peek()[0] == "close":
                            args.append((None, None))
                    continue
            args.append((name, self.expr(5)))
            nxt = self.peek()
            if nxt[:2] == ("op", ","):
                self.next()
                if self.peek()[0] == "close":
                    args.append((None, None))
            elif nxt[0] != "close":
                self.err("',' or closing bracket expected in an argument list", nxt)
        want = {"(": ")", "[": "]", "[[": "]"}[text]
        self.expect("close", want)
        if text == "[[":
            if self.t[self.i][:2] != ("close", "]"):     # the two ']' of ']]' are adjacent tokens
                self.err("']]' expected")
            self.i += 1
        self.ctx.pop()
        if text == "(":
            return Node("call", line, left, args)
        return Node("index", line, text, left, args)

    def _nl_ahead(self, j):
        k = 0
        while self.t[j + k][0] == "nl":
            k += 1
        return k


def parse(src):
    """The top-level expressions of an R source text; SyntaxError (with the line) when R would not parse it."""
    return Parser(src).program()


# ---------------------------------------------------------------------------------------------------------------------
# analyses
# ---------------------------------------------------------------------------------------------------------------------
def walk(node):
    """Every Node below (and including) node."""
    if isinstance(node, Node):
        yield node
        for x in node.a:
            yield from walk(x)
    elif isinstance(node, (list, tuple)):
        for x in node:
            yield from walk(x)


def top_level_definitions(prog):
    """{name: value node} of `name <- value` / `name = value` at top level."""
    out = {}
    for e in prog:
        if e.kind == "binop" and e.a[0] in ("<-", "=", "<<-") and e.a[1].kind in ("name", "str"):
            out[e.a[1].a[0]] = e.a[2]
    return out


def _assign_target(lhs):
    """The variable an assignment (re)binds: x, x$a, x[i], x@s, f(x) -> x; the replacement function's name or None."""
    repl = None
    while True:
        if lhs.kind in ("name", "str"):
            return lhs.a[0], repl
        if lhs.kind in ("dollar", "at"):
            lhs = lhs.a[0]
        elif lhs.kind == "index":
            lhs = lhs.a[1]
        elif lhs.kind == "paren":
            lhs = lhs.a[0]
        elif lhs.kind == "call" and lhs.a[1] and lhs.a[1][0][1] is not None:
            if repl is None and lhs.a[0].kind == "name":
                repl = lhs.a[0].a[0]
            lhs = lhs.a[1][0][1]
        else:
            return None, repl


def local_names(fn):
    """Names a function binds in its own frame: formals, `<-` / `=` / `->` targets, loop variables (not those of nested
    functions)."""
    names = {f for f, _ in fn.a[0]}

    def visit(n):
        if isinstance(n, (list, tuple)):
            for x in n:
                visit(x)
            return
        if not isinstance(n, Node):
            return
        if n.kind == "function":
            for _, d in n.a[0]:
                visit(d)
            return                                   # its body has its own frame
        if n.kind == "binop" and n.a[0] in ("<-", "="):
            tgt, _ = _assign_target(n.a[1])
            if tgt:
                names.add(tgt)
        elif n.kind == "binop" and n.a[0] == "->":
            tgt, _ = _assign_target(n.a[2])
            if tgt:
                names.add(tgt)
        elif n.kind == "for":
            names.add(n.a[0])
        for x in n.a:
            visit(x)

    visit(fn.a[1])
    return names


def free_variables(fn, enclosing=frozenset()):
    """{name: first line} of the variables a function (and its nested functions) read but do not bind."""
    bound = set(enclosing) | local_names(fn)
    free = {}

    def use(name, line):
        if name not in bound and name not in free:
            free[name] = line

    def visit(n):
        if isinstance(n, (list, tuple)):
            for x in n:
                visit(x)
            return
        if not isinstance(n, Node):
            return
        k = n.kind
        if k == "name":
            use(n.a[0], n.line)
        elif k == "function":
            for name, line in free_variables(n, bound).items():
                use(name, line)
        elif k in ("dollar", "at"):
            visit(n.a[0])
            if isinstance(n.a[1], Node):
                visit(n.a[1])
        elif k == "ns":
            pass
        elif k == "call":
            visit(n.a[0])
            for _, v in n.a[1]:
                visit(v)
        elif k == "index":
            visit(n.a[1])
            for _, v in n.a[2]:
                visit(v)
        elif k == "for":
            visit(n.a[1])
            visit(n.a[2])
        elif k == "binop" and n.a[0] in ("<-", "=", "<<-") and n.a[1].kind == "call":
            # f(x, i) <- v: reads x and i, calls `f<-`
            _, repl = _assign_target(n.a[1])
            for _, v in n.a[1].a[1]:
                visit(v)
            if n.a[1].a[0].kind != "name":
                visit(n.a[1].a[0])
            elif repl:
                use(repl + "<-", n.line)
            visit(n.a[2])
        else:
            for x in n.a:
                visit(x)

    for _, d in fn.a[0]:
        visit(d)
    visit(fn.a[1])
    return free


def calls_of(node):
    """Every call `name(...)` / `pkg::name(...)` below node: (callee name, [(argument name or None, value)], line)."""
    for n in walk(node):
        if n.kind == "call" and n.a[0].kind in ("name", "ns"):
            yield (n.a[0].a[0] if n.a[0].kind == "name" else n.a[0].a[1]), n.a[1], n.line


def check_call(callee, args, formals):
    """R's argument matching against a formals list [(name, has_default)]: returns a list of complaints (exact names
    only -- partial matching is legal R but never intended in package code)."""
    names = [f for f, _ in formals]
    dots = "..." in names
    problems = []
    named = [a for a, v in args if a is not None]
    for a in named:
        if a not in names and not dots:
            problems.append(f"{callee}(): unknown argument {a!r}")
    if len(set(named)) != len(named):
        problems.append(f"{callee}(): argument given twice")
    positional = sum(1 for a, v in args if a is None and v is not None)
    free_slots = [f for f in names if f not in named and f != "..."]
    if positional > len(free_slots) and not dots:
        problems.append(f"{callee}(): {positional} positional arguments for {len(free_slots)} remaining formals")
    supplied = set(named) | set(free_slots[:positional])
    for f, has_default in formals:
        if f != "..." and not has_default and f not in supplied:
            problems.append(f"{callee}(): argument {f!r} has no default and is not supplied")
    return problems


def element_reads(node):
    """(name, line) of every x$name / x@name below node that is READ (not the target of an assignment)."""
    written = set()
    for n in walk(node):
        if n.kind == "binop" and n.a[0] in ("<-", "=", "<<-"):
            lhs = n.a[1]                                         # rownames(x$a) <- v still reads x$a
            if lhs.kind in ("dollar", "at"):
                written.add(id(lhs))
    for n in walk(node):
        if n.kind in ("dollar", "at") and id(n) not in written and isinstance(n.a[1], str):
            yield n.kind, n.a[1], n.line


def element_writes(node):
    """Names set as list(name = ...) / c(name = ...) elements, x$name <- v / x@name <- v targets, attr(x, "name")."""
    out = set()
    for n in walk(node):
        if n.kind == "call":
            for a, _ in n.a[1]:
                if a is not None:
                    out.add(a)
        if n.kind == "binop" and n.a[0] in ("<-", "=", "<<-") and n.a[1].kind in ("dollar", "at") \
                and isinstance(n.a[1].a[1], str):
            out.add(n.a[1].a[1])
    return out
