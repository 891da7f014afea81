"""Convert the reference's test fixtures (DATA, not code) into one small .npz that travels with the
repo, because /root/reference does not exist on the GPU box.

    python tests/golden/make_fixtures.py          # run in the build container; writes fixtures.npz

Sources (read-only): /root/reference/test/{thing,test,randlap,onetoall,ref_R,ref_S_test}.jl — literal
``SparseMatrixCSC(m, n, colptr, rowval, nzval)`` constructors (1-based Int64); ref_split_test.txt;
lin_elastic_2d.jld2 and bug.jld2 (JLD2 = HDF5 subset, h5py is absent: raw offsets).
Arrays are stored as found (1-based indices); tests/fixtures.py subtracts the offset.
"""
import os
import re
import struct

import numpy as np

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.npz")


def parse_jl(path):
    s = open(path).read()
    m, n = map(int, re.search(r"Gm, Gn = (\d+), (\d+)", s).groups())

    def arr(name, dtype):
        body = re.search(name + r"\s*=\s*\[(.*?)\]", s, re.S).group(1)
        return np.array([t for t in body.replace("\n", " ").split(",") if t.strip()], dtype=dtype)

    return m, n, arr("Gcolptr", np.int64), arr("Growval", np.int64), arr("Gnzval", np.float64)


def main():
    out = {}
    for name in ["thing", "test", "randlap", "onetoall", "ref_R", "ref_S_test"]:
        m, n, cp, rv, nz = parse_jl(os.path.join(REF, name + ".jl"))
        out[name + "_shape"] = np.array([m, n])
        out[name + "_colptr"], out[name + "_rowval"], out[name + "_nzval"] = cp, rv, nz
    out["ref_split"] = np.loadtxt(os.path.join(REF, "ref_split_test.txt")).astype(np.int64)

    raw = open(os.path.join(REF, "lin_elastic_2d.jld2"), "rb").read()
    base = 512
    out["elastic_shape"] = np.array([208, 208])
    out["elastic_colptr"] = np.frombuffer(raw, dtype="<i8", count=209, offset=base + 0x1310).copy()
    out["elastic_rowval"] = np.frombuffer(raw, dtype="<i8", count=2632, offset=base + 0x19F8).copy()
    out["elastic_nzval"] = np.frombuffer(raw, dtype="<f8", count=2632, offset=base + 0x6CA0).copy()
    out["elastic_b"] = np.frombuffer(raw, dtype="<f8", count=208, offset=base + 0xBF48).copy()
    out["elastic_B"] = np.frombuffer(raw, dtype="<f8", count=624, offset=base + 0xC638).copy().reshape(208, 3, order="F")
    assert out["elastic_colptr"][0] == 1 and out["elastic_colptr"][-1] == 2633
    assert out["elastic_rowval"].min() == 1 and out["elastic_rowval"].max() == 208

    raw = open(os.path.join(REF, "bug.jld2"), "rb").read()
    out["bug_shape"] = np.array([4, 4])
    out["bug_colptr"] = np.frombuffer(raw, dtype="<i8", count=5, offset=5291).copy()
    out["bug_rowval"] = np.frombuffer(raw, dtype="<i8", count=16, offset=5388).copy()
    out["bug_nzval"] = np.frombuffer(raw, dtype="<f8", count=16, offset=5581).copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    for k in ("bug_colptr", "bug_rowval", "bug_nzval"):
        print(k, out[k])
    assert list(out["bug_colptr"]) == [1, 5, 9, 13, 17]


if __name__ == "__main__":
    main()
