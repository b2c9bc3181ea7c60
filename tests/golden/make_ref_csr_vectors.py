#!/usr/bin/env python
"""Generate tests/golden/ref_csr_vectors.npz from the REFERENCE's own la/matrix_csr_impl.h.

The reference header is compiled in place from /root/reference into oracle/_ref/libref_csr.so
(oracle/Makefile, target `ref`; the sources are never copied).  This script runs that library on seeded
inputs and stores inputs + outputs, so that the oracle (and anything checked against the oracle) stays
pinned to the reference's arithmetic on machines where /root/reference and oracle/_ref do not exist.

    python tests/golden/make_ref_csr_vectors.py        # needs oracle/_ref/libref_csr.so

Cases (per block size (bs0, bs1) in BLOCK_SIZES), following la/matrix_csr_impl.h:
  insert_csr            :67-109   set and add of dense blocks into random rows of a random pattern
  insert_blocked_csr    :112-156  the same through the blocked entry point (kind 1)
  insert_nonblocked_csr :159-201  scalar rows / columns into a blocked matrix
  spmv / spmvT          :204-281  y += A x and y += A^T x
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "..", "..", "oracle", "_ref", "libref_csr.so")
BLOCK_SIZES = [(1, 1), (2, 2), (3, 3), (2, 3)]
NROW, NCOL, SEED = 16, 20, 20261017


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pattern(rng):
    dense = rng.random((NROW, NCOL)) < 0.3
    dense[np.arange(NROW), np.arange(NROW)] = True
    row_ptr = np.concatenate([[0], np.cumsum(dense.sum(1))]).astype(np.int64)
    cols = np.concatenate([np.flatnonzero(dense[i]) for i in range(NROW)]).astype(np.int32)
    return dense, row_ptr, cols


def main():
    ref = C.CDLL(REF_SO)
    ref.ref_spmv.restype = None
    out = {}
    rng = np.random.default_rng(SEED)
    dense, row_ptr, cols = pattern(rng)
    out["row_ptr"], out["cols"] = row_ptr, cols
    for bs0, bs1 in BLOCK_SIZES:
        tag = f"bs{bs0}{bs1}"
        bs2 = bs0 * bs1
        # ---- a sequence of inserts applied to ONE array (set, then adds): the final array is the fixture
        # (kind 1, insert_blocked_csr: blocked indices into an UNBLOCKED matrix - the scalar expansion of the pattern)
        dense_s = np.kron(dense.astype(np.int8), np.ones((bs0, bs1), dtype=np.int8)).astype(bool)
        row_ptr_s = np.concatenate([[0], np.cumsum(dense_s.sum(1))]).astype(np.int64)
        cols_s = np.concatenate([np.flatnonzero(dense_s[i]) for i in range(NROW * bs0)]).astype(np.int32)
        out[f"{tag}_scalar_row_ptr"], out[f"{tag}_scalar_cols"] = row_ptr_s, cols_s
        for kind, name in ((0, "csr"), (1, "blocked")):
            rp_k, cols_k = (row_ptr, cols) if kind == 0 else (row_ptr_s, cols_s)
            data = rng.random(len(cols) * bs2)
            out[f"{tag}_{name}_data0"] = data.copy()
            ops = []
            for trial in range(12):
                rows = rng.choice(NROW, size=2, replace=False).astype(np.int32)
                common = np.flatnonzero(dense[rows[0]] & dense[rows[1]])
                if len(common) == 0:
                    rows = rows[:1]
                    common = np.flatnonzero(dense[rows[0]])
                xc = rng.choice(common, size=min(3, len(common)), replace=False).astype(np.int32)
                x = rng.random(len(rows) * bs0 * len(xc) * bs1)
                op = int(trial % 3 != 0)  # set, add, add, ...
                err = ref.ref_insert(kind, bs0, bs1, ptr(data), C.c_size_t(len(data)), ptr(cols_k), C.c_size_t(len(cols_k)),
                                     ptr(rp_k), C.c_size_t(len(rp_k)), ptr(x), ptr(rows), len(rows), ptr(xc), len(xc), op)
                assert err == 0, err
                ops.append((rows, xc, x, op))
            out[f"{tag}_{name}_nops"] = np.array(len(ops))
            for k, (rows, xc, x, op) in enumerate(ops):
                out[f"{tag}_{name}_op{k}_rows"], out[f"{tag}_{name}_op{k}_cols"] = rows, xc
                out[f"{tag}_{name}_op{k}_x"], out[f"{tag}_{name}_op{k}_add"] = x, np.array(op)
            out[f"{tag}_{name}_data1"] = data
        # ---- non-blocked insert: scalar rows / columns of the blocked matrix
        data = rng.random(len(cols) * bs2)
        out[f"{tag}_nonblocked_data0"] = data.copy()
        i = int(rng.integers(NROW))
        avail = np.flatnonzero(dense[i])
        xr = np.array([i * bs0 + int(rng.integers(bs0))], dtype=np.int32)
        xc = np.array([int(avail[0]) * bs1 + int(rng.integers(bs1)), int(avail[-1]) * bs1], dtype=np.int32)
        x = rng.random(len(xr) * len(xc))
        err = ref.ref_insert_nonblocked(bs0, bs1, ptr(data), C.c_size_t(len(data)), ptr(cols), C.c_size_t(len(cols)),
                                        ptr(row_ptr), C.c_size_t(len(row_ptr)), ptr(x), ptr(xr), len(xr), ptr(xc), len(xc), 1)
        assert err == 0, err
        out[f"{tag}_nonblocked_rows"], out[f"{tag}_nonblocked_cols"], out[f"{tag}_nonblocked_x"] = xr, xc, x
        out[f"{tag}_nonblocked_data1"] = data
        # ---- spmv / spmvT
        vals = rng.random(len(cols) * bs2)
        x = rng.random(NCOL * bs1)
        y = rng.random(NROW * bs0)
        out[f"{tag}_spmv_vals"], out[f"{tag}_spmv_x"], out[f"{tag}_spmv_y0"] = vals, x, y.copy()
        rb, re = np.ascontiguousarray(row_ptr[:-1]), np.ascontiguousarray(row_ptr[1:])
        ref.ref_spmv(0, ptr(vals), C.c_size_t(len(vals)), ptr(rb), ptr(re), C.c_size_t(NROW), ptr(cols), C.c_size_t(len(cols)),
                     ptr(x), C.c_size_t(len(x)), ptr(y), C.c_size_t(len(y)), bs0, bs1)
        out[f"{tag}_spmv_y1"] = y
        xt = rng.random(NROW * bs0)
        yt = rng.random(NCOL * bs1)
        out[f"{tag}_spmvT_x"], out[f"{tag}_spmvT_y0"] = xt, yt.copy()
        ref.ref_spmv(1, ptr(vals), C.c_size_t(len(vals)), ptr(rb), ptr(re), C.c_size_t(NROW), ptr(cols), C.c_size_t(len(cols)),
                     ptr(xt), C.c_size_t(len(xt)), ptr(yt), C.c_size_t(len(yt)), bs0, bs1)
        out[f"{tag}_spmvT_y1"] = yt
    path = os.path.join(HERE, "ref_csr_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
