"""Full-size parity against the oracle on a SAMPLE of matrix rows (checker code: used by tests/test_gpu_fullsize.py
and by the ``parity`` leg of bench.py, never by the product).

All cells incident to the chosen (block) rows are cut out of the big mesh, renumbered compactly and assembled by the CPU
oracle; the complete rows (column indices bit-exact, values row-scaled) are compared with the same rows of the GPU matrix.
"""

import numpy as np

from tests import problems as P


def oracle_rows(torch, oracle, pb, kid, consts, markers_host, rows):
    """Rows `rows` (block indices, owned, complete on this rank) of the matrix by the CPU oracle on the sub-mesh of
    their incident cells.  Returns (global ids of the sub-mesh dofs, oracle pattern, oracle values, cells used)."""
    V, msh = pb["V"], pb["mesh"]
    dm = V.dofmap.dev
    rows_dev = torch.from_numpy(rows).to(dm.device)
    sel = torch.zeros(dm.shape[0], dtype=torch.bool, device=dm.device)
    lut = torch.zeros(int(dm.max()) + 1, dtype=torch.bool, device=dm.device)
    lut[rows_dev.long()] = True
    for i in range(dm.shape[1]):
        sel |= lut[dm[:, i].long()]
    cells = torch.nonzero(sel).reshape(-1)
    dm_s = dm[cells].cpu().numpy()
    xd_s = msh.x_dofmap[cells].cpu().numpy()
    ud, inv = np.unique(dm_s, return_inverse=True)
    ux, invx = np.unique(xd_s, return_inverse=True)
    x_l = msh.x[torch.from_numpy(ux).to(dm.device).long()].cpu().numpy()
    bs = pb["bs"]
    p = P.Problem(x_l, invx.reshape(xd_s.shape).astype(np.int32), inv.reshape(dm_s.shape).astype(np.int32), len(ud), bs,
                  msh.cell_type)
    bc_l = None
    if markers_host is not None:
        bc_l = np.ascontiguousarray(markers_host.reshape(-1, bs)[ud].reshape(-1))
    pat, ref = P.oracle_assemble_matrix(oracle, p, kid, constants=consts, bc=bc_l)
    if bc_l is not None:
        loc = np.searchsorted(ud, rows).astype(np.int32)
        unrolled = oracle.unroll_dofs(loc, bs)
        unrolled = unrolled[bc_l[unrolled] != 0]
        oracle.set_diagonal(ref, pat.edges, pat.offsets, bs, bs, unrolled.astype(np.int32), 1.0)
    return ud, pat, ref, int(cells.numel())


def compare_rows(A, rows, ud, pat, ref, bs):
    """max over the sampled rows of |a_ij - ref_ij| / max_k |ref_ik| (every scalar row of a block row scaled by its
    own maximum); the column indices must agree exactly."""
    indptr = A.indptr
    indices = A.indices
    bs2 = bs * bs
    worst = 0.0
    data = A._values()
    for r in rows:
        l = int(np.searchsorted(ud, r))
        o0, o1 = int(pat.offsets[l]), int(pat.offsets[l + 1])
        g0, g1 = int(indptr[r]), int(indptr[r + 1])
        assert np.array_equal(ud[pat.edges[o0:o1]], indices[g0:g1]), f"columns of row {r} differ"
        ref_r = ref[o0 * bs2:o1 * bs2]
        got = data[g0 * bs2:g1 * bs2].cpu().numpy()
        rr = ref_r.reshape(o1 - o0, bs, bs)
        gg = got.reshape(o1 - o0, bs, bs)
        for i in range(bs):
            scale = np.abs(rr[:, i, :]).max()
            scale = scale if scale > 0 else 1.0
            worst = max(worst, float(np.abs(gg[:, i, :] - rr[:, i, :]).max() / scale))
    return worst
