"""A small lexer for R source (test infrastructure).  R is not installed in this image, so r_shim/R/lpb200_wrappers.R
cannot be parsed by R itself; this checks what a never-run wrapper file typically gets wrong: unterminated strings,
brackets that do not nest, .Call()s whose routine or argument count does not match the shim's registration table, and
list element names the C side looks up but the R side never sets."""
import re


def tokens(src):
    """Yields (kind, text, line): kind in {'str', 'name', 'num', 'op', 'open', 'close'}; comments and blanks dropped."""
    i, n, line = 0, len(src), 1
    while i < n:
        c = src[i]
        if c == "\n":
            line += 1
            i += 1
        elif c in " \t\r":
            i += 1
        elif c == "#":
            while i < n and src[i] != "\n":
                i += 1
        elif c in "\"'`":
            j = i + 1
            while j < n and src[j] != c:
                if src[j] == "\\" and c != "`":
                    j += 1
                if src[j] == "\n":
                    line += 1
                j += 1
            if j >= n:
                raise SyntaxError(f"line {line}: unterminated {c} string")
            yield ("name" if c == "`" else "str", src[i + 1:j], line)
            i = j + 1
        elif c.isalpha() or c == "." and i + 1 < n and (src[i + 1].isalpha() or src[i + 1] in "._"):
            j = i
            while j < n and (src[j].isalnum() or src[j] in "._"):
                j += 1
            yield ("name", src[i:j], line)
            i = j
        elif c.isdigit() or c == "." and i + 1 < n and src[i + 1].isdigit():
            m = re.match(r"(0[xX][0-9a-fA-F]+|\d*\.?\d*([eE][+-]?\d+)?)[Li]?", src[i:])
            yield ("num", m.group(0), line)
            i += max(1, len(m.group(0)))
        elif src.startswith("[[", i):
            yield ("open", "[[", line)
            i += 2
        elif c in "([{":
            yield ("open", c, line)
            i += 1
        elif c in ")]}":
            yield ("close", c, line)
            i += 1
        else:
            m = re.match(r"%[^%\n]*%|<<-|->>|<-|->|<=|>=|==|!=|&&|\|\||::|:::|[-+*/^<>=!&|~?:$@,;\\]", src[i:])
            if not m:
                raise SyntaxError(f"line {line}: unexpected character {c!r}")
            yield ("op", m.group(0), line)
            i += len(m.group(0))


def check_nesting(src):
    """Brackets nest properly ('[[' closes with two ']').  Returns the token list."""
    toks = list(tokens(src))
    stack = []
    for kind, text, line in toks:
        if kind == "open":
            stack.append((text, line))
        elif kind == "close":
            if not stack:
                raise SyntaxError(f"line {line}: unmatched {text}")
            top, l0 = stack[-1]
            want = {"(": ")", "{": "}", "[": "]", "[[": "]"}[top]
            if text != want:
                raise SyntaxError(f"line {line}: {text} closes {top} opened on line {l0}")
            if top == "[[":
                stack[-1] = ("[", l0)     # the first ']' of ']]'
            else:
                stack.pop()
    if stack:
        raise SyntaxError(f"line {stack[-1][1]}: {stack[-1][0]} never closed")
    return toks


def calls(toks, fname):
    """Every call fname(...): list of (line, [argument token lists]) split at top-level commas."""
    out = []
    for k, (kind, text, line) in enumerate(toks):
        if kind == "name" and text == fname and k + 1 < len(toks) and toks[k + 1][:2] == ("open", "("):
            depth, args, cur = 0, [], []
            for t in toks[k + 1:]:
                if t[0] == "open":
                    depth += 1 if t[1] != "[[" else 2
                    if depth == 1:
                        continue
                elif t[0] == "close":
                    depth -= 1
                    if depth == 0:
                        break
                if depth == 1 and t[:2] == ("op", ","):
                    args.append(cur)
                    cur = []
                else:
                    cur.append(t)
            if cur or args:
                args.append(cur)
            out.append((line, args))
    return out


def function_defs(toks):
    """Top-level `name <- function(args)` definitions: {name: [formal names]}."""
    out = {}
    for k in range(len(toks) - 3):
        if toks[k][0] == "name" and toks[k + 1][:2] == ("op", "<-") and toks[k + 2][:2] == ("name", "function"):
            depth, formals, expect = 0, [], True
            for t in toks[k + 3:]:
                if t[0] == "open":
                    depth += 1
                    continue
                if t[0] == "close":
                    depth -= 1
                    if depth == 0:
                        break
                    continue
                if depth == 1:
                    if t[:2] == ("op", ","):
                        expect = True
                    elif expect and t[0] == "name":
                        formals.append(t[1])
                        expect = False
            out[toks[k][1]] = formals
    return out
