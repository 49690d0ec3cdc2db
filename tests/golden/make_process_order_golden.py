"""Generates tests/golden/process_sc_data_order.json: the reference's processSCData (processSCData.R:85-365) EXECUTED from its
source text by tests/r_eval.py on count matrices whose columns are ordered differently from dfAnnotation$cell (run in the build
container, where /root/reference exists):

    python -m tests.golden.make_process_order_golden

The reference sorts the count columns by match(colnames(counts), dfAnnotation$cell) (:255-258) -- the inverse of the permutation
that would line them up with the annotation -- and pairs columns and annotation rows by position afterwards.  The JSON records
what it returns (column names and counts of matCountsProc, dfAnnotationProc, the size factors looked up by column name) for a
swap (its own inverse: lined up) and for a 3-cycle (not lined up); tests/test_host_logic.py holds api.processSCData to it."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden.make_reference_exec_golden import B, D, S, Reference  # noqa: E402
from tests.r_eval import RList, RVec, as_py  # noqa: E402


def run(ref, X, genes, colnames, cells, t, sf):
    G, N = X.shape
    mat = RVec(np.asarray(X, dtype=np.float64).reshape(-1, order="F"), dim=(G, N), dimnames=(genes, colnames))
    df = RList([S(cells), D(t)], ["cell", "continuous"])
    df.attrs = {"class": "data.frame", "row.names": list(cells)}
    res = ref.call("processSCData", counts=mat, dfAnnotation=df, vecConfoundersDisp=None, vecConfoundersMu=None,
                   matPiConstPredictors=None, vecNormConstExternal=RVec(np.asarray(sf, dtype=np.float64), names=list(cells)),
                   strDispModelFull=S("constant"), strDispModelRed=S("constant"), strMuModel=S("splines"), scaDFSplinesDisp=D(3),
                   scaDFSplinesMu=D(3), scaMaxEstimationCycles=D(20), boolVerbose=B(False), boolSuperVerbose=B(False))
    obj = res.get("objLP")
    slot = lambda name: obj.v[obj.names.index(name)]     # noqa: E731
    m, a = slot("matCountsProc"), slot("dfAnnotationProc")
    return dict(colnames=list(m.dimnames[1]), rownames=list(m.dimnames[0]), counts=m.mat().tolist(),
                annotation_cell=list(a.v[a.names.index("cell")].v), annotation_continuous=as_py(a.v[a.names.index("continuous")]),
                vecNormConstExternalProc=as_py(res.get("vecNormConstExternalProc")),
                report=slot("strReport").v[0].split("\n")[1:])


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    ref = Reference(root)
    rng = np.random.default_rng(8)
    cells = [f"cell_{j + 1}" for j in range(6)]
    genes = [f"gene_{i + 1}" for i in range(4)]
    X = rng.poisson(2.0, size=(4, 6)) + 1
    X[1, :] = 0                                  # an all-zero gene
    t = [0.0, 0.2, float("nan"), 0.6, 0.8, 1.0]    # annotation row 3 has no pseudotime
    sf = [0.5, 0.75, 1.0, 1.25, 1.5, 2.0]         # named by annotation cell
    cases = []
    for name, order in (("same order", [0, 1, 2, 3, 4, 5]), ("one swap", [1, 0, 2, 3, 4, 5]), ("two swaps", [1, 0, 2, 5, 4, 3]),
                        ("3-cycle", [1, 2, 0, 3, 4, 5]), ("reversed", [5, 4, 3, 2, 1, 0]), ("6-cycle", [1, 2, 3, 4, 5, 0])):
        colnames = [cells[k] for k in order]     # counts column j holds cell order[j]; X column j belongs to that cell
        cases.append(dict(name=name, counts=X.tolist(), genes=genes, colnames=colnames, cells=cells, t=[None if v != v else v for v in t],
                          sf=sf, reference=run(ref, X, genes, colnames, cells, t, sf)))
        print(name, cases[-1]["reference"]["colnames"], cases[-1]["reference"]["annotation_cell"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "process_sc_data_order.json")
    json.dump(dict(note="processSCData of the reference executed by tests/r_eval.py; see make_process_order_golden.py", cases=cases),
              open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
