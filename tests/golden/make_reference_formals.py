"""Generates tests/golden/reference_formals.json from the reference's R sources (run in the build container, where
/root/reference exists; the GPU box and the tests only read the JSON):
    python tests/golden/make_reference_formals.py [/root/reference]
For every function of the reference's API that lineagepulse_b200.api mirrors: the formal argument names in order and,
where the default is a single literal (number, string, TRUE / FALSE / NULL), that default.
Also writes tests/golden/reference_returns.json (element names of the lists / data frames the boundary functions return)
and tests/golden/reference_messages.json: the string literals of every stop() / warning() call of the reference
(file, line, the literals in order) -- the error behaviour the host mirror keeps."""
import glob
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import r_lint  # noqa: E402

FUNCTIONS = ["runLineagePulse", "processSCData", "calcNormConst", "fitLPModels", "fitModel", "fitMuDisp", "fitPi",
             "evalLogLikMatrix", "runDEAnalysis", "testDropout", "calcPostDrop_Matrix", "getPostDrop", "getNormData",
             "getFitsMean", "getFitsDispersion", "getFitsDropout", "simulateContinuousDataSet", "initMuModel",
             "initDispModel", "initDropModel", "evalLogLikZINB", "evalLogLikNB", "evalImpulseModel", "evalDropoutModel"]


def formals_with_defaults(toks, name):
    for k in range(len(toks) - 3):
        if toks[k][:2] == ("name", name) and toks[k + 1][:2] == ("op", "<-") and toks[k + 2][:2] == ("name", "function"):
            depth, out, cur = 0, [], []
            for t in toks[k + 3:]:
                if t[0] == "open":
                    depth += 1
                    if depth == 1:
                        continue
                elif t[0] == "close":
                    depth -= 1
                    if depth == 0:
                        break
                if depth == 1 and t[:2] == ("op", ","):
                    out.append(cur)
                    cur = []
                else:
                    cur.append(t)
            if cur:
                out.append(cur)
            res = []
            for arg in out:
                default = None
                if len(arg) == 3 and arg[1][:2] == ("op", "=") and arg[2][0] in ("num", "str", "name"):
                    kind, text = arg[2][:2]
                    if kind == "str":
                        default = {"str": text}
                    elif kind == "num":
                        default = {"num": float(text.rstrip("L"))}
                    elif text in ("TRUE", "FALSE", "NULL"):
                        default = {"lit": text}
                res.append([arg[0][1], default])
            return res
    return None


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    toks = []
    for f in sorted(glob.glob(os.path.join(root, "R", "*.R"))):
        toks += list(r_lint.tokens(open(f).read()))
    out = {}
    for fn in FUNCTIONS:
        fm = formals_with_defaults(toks, fn)
        assert fm is not None, fn
        out[fn] = fm
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_formals.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, {k: len(v) for k, v in out.items()})
    msgs = []
    for f in sorted(glob.glob(os.path.join(root, "R", "*.R"))):
        ftoks = list(r_lint.tokens(open(f).read()))
        for fn in ("stop", "warning"):
            for line, args in r_lint.calls(ftoks, fn):
                lits = [t[1] for a in args for t in a if t[0] == "str"]
                if lits:
                    msgs.append({"file": os.path.basename(f), "line": line, "call": fn, "literals": lits})
    msgs.sort(key=lambda m: (m["file"], m["line"]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_messages.json")
    json.dump(msgs, open(path, "w"), indent=1)
    print("wrote", path, len(msgs), "messages")



# ---- names of what the boundary functions return (appended: reference_returns.json) --------------------------------
def body_tokens(toks, name):
    """Tokens of the body { ... } of the top-level definition `name <- function(...) { ... }`."""
    for k in range(len(toks) - 3):
        if toks[k][:2] == ("name", name) and toks[k + 1][:2] == ("op", "<-") and toks[k + 2][:2] == ("name", "function"):
            depth, j = 0, k + 3
            while True:            # skip the formals
                if toks[j][0] == "open":
                    depth += 1
                elif toks[j][0] == "close":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            j += 1
            assert toks[j][:2] == ("open", "{"), name
            start, depth = j, 0
            while True:
                if toks[j][0] == "open":
                    depth += 1 if toks[j][1] != "[[" else 2
                elif toks[j][0] == "close":
                    depth -= 1
                    if depth == 0:
                        return toks[start:j + 1]
                j += 1
    return None


def named_args(args):
    return [a[0][1] for a in args if len(a) >= 3 and a[0][0] in ("name", "str") and a[1][:2] == ("op", "=")]


def returns_main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    toks = []
    for f in sorted(glob.glob(os.path.join(root, "R", "*.R"))):
        toks += list(r_lint.tokens(open(f).read()))
    out = {}
    for fn in ("fitMuDisp", "fitPi", "fitModel", "fitMuDispGene", "fitMuDispGeneImpulse", "fitPi_SingleCell", "fitPi_ManyCells",
               "processSCData", "simulateContinuousDataSet"):
        body = body_tokens(toks, fn)
        lists = [named_args(a) for _, a in r_lint.calls(body, "list")]
        # the returned list: the last list(...) call with named elements inside return(...)
        rets = []
        for line, args in r_lint.calls(body, "return"):
            inner = [t for a in args for t in a]
            for _, la in r_lint.calls(inner, "list"):
                if named_args(la):
                    rets.append(named_args(la))
        out[fn] = rets[-1] if rets else None
    for fn in ("runDEAnalysis", "testDropout"):
        body = body_tokens(toks, fn)
        frames = [named_args(a) for _, a in r_lint.calls(body, "data.frame")]
        out[fn] = max(frames, key=len)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_returns.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
    for k, v in out.items():
        print(" ", k, v)


# ---- every name the R wrappers may refer to (reference_names.json) ---------------------------------------------------
def names_main():
    """All top-level functions of the reference with their formals [(name, has_default)], the slots of its S4 class and
    every element name it sets (list(name = ), x$name <- ) or reads (x$name): what r_shim/R/lpb200_wrappers.R is checked
    against by tests/test_host_logic.py::test_r_wrappers_semantic_checks (parsed with tests/r_parse.py)."""
    from tests import r_parse
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    functions, elements, slots = {}, set(), set()
    for f in sorted(glob.glob(os.path.join(root, "R", "*.R"))):
        prog = r_parse.parse(open(f).read())
        for name, val in r_parse.top_level_definitions(prog).items():
            if val.kind == "function":
                functions[name] = [[a, d is not None] for a, d in val.a[0]]
        elements |= r_parse.element_writes(prog)
        elements |= {name for kind, name, _ in r_parse.element_reads(prog) if kind == "dollar"}
        for callee, args, _ in r_parse.calls_of(prog):
            if callee == "setClass":
                for a, v in args:
                    if a in ("slots", "representation") and v is not None and v.kind == "call":
                        slots |= {n for n, _ in v.a[1] if n}
    assert "LineagePulseObject" and "matCountsProc" in slots and "fitModel" in functions
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_names.json")
    json.dump({"functions": functions, "slots": sorted(slots), "elements": sorted(elements)}, open(path, "w"),
              indent=0, sort_keys=True)
    print("wrote", path, len(functions), "functions,", len(slots), "slots,", len(elements), "element names")


if __name__ == "__main__":
    main()
    returns_main()
    names_main()
