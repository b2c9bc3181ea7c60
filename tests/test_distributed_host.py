"""N>1 path on CPU: (1) the oracle's simulated-rank restatement reproduces serial results;
(2) the product's host plan construction (IndexMap, Scatterer, SparsityPattern.finalize,
MatrixCSR ghost plan), run as two gloo processes, is bit-identical to the oracle's plans."""

import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from dolfinx_b200 import mesh as M  # noqa: E402
from tests import problems as P  # noqa: E402


def brick_inputs(nranks, nglob, cell="tet"):
    pg = M.pgrid_for(nranks)
    parts = [M.BoxPartition(r, pg, nglob) for r in range(nranks)]
    out = []
    for part in parts:
        x = part.coordinates()
        xd, dm, n_owned, ghosts, owners = part.tets() if cell == "tet" else part.hexes()
        out.append(dict(x=x, x_dofmap=xd, dofmap=dm, n_owned=n_owned, ghosts=ghosts, owners=owners, part=part))
    return out


def oracle_world(O, inputs, bs=1):
    maps = O.make_index_maps([i["n_owned"] for i in inputs], [i["ghosts"] for i in inputs], [i["owners"] for i in inputs])
    rows, cols = [], []
    for i in inputs:
        r, c = O.sparsity_insert_cells(np.arange(len(i["dofmap"])), i["dofmap"], i["dofmap"])
        rows.append(r)
        cols.append(c)
    pats = O.sparsity_finalize(maps, maps, (bs, bs), rows, cols)
    mats = O.make_matrices(pats)
    return maps, pats, mats


@pytest.mark.parametrize("nranks,nglob", [(2, (4, 3, 3)), (3, (6, 2, 3)), (4, (4, 4, 2)), (8, (4, 4, 4))])
def test_oracle_distributed_assembly_equals_serial(oracle, nranks, nglob):
    """Assemble per rank + scatter_rev, gather by global index: identical structure, values to 1e-12,
    of the serial assembly (cf. cpp/test/matrix.cpp:59-64, serial vs parallel norm)."""
    O = oracle
    inputs = brick_inputs(nranks, nglob)
    maps, pats, mats = oracle_world(O, inputs)
    kappa = np.array([2.0])
    for i, A in zip(inputs, mats):
        O.assemble_matrix(O.K_POISSON_P1_TET_A, i["x_dofmap"], i["x"], np.arange(len(i["dofmap"])), i["dofmap"], 1,
                          i["dofmap"], 1, A.data, A.cols, A.row_ptr, constants=kappa)
    total_before = sum(A.data.sum() for A in mats)
    O.matrix_scatter_rev(mats)
    # python/test/unit/la/test_matrix_csr.py:111-147: sum preserved, ghost rows zeroed
    assert sum(A.data.sum() for A in mats) == pytest.approx(total_before, abs=1e-10)
    for A in mats:
        assert np.all(A.data[A.row_ptr[A.index_maps[0].size_local]:] == 0)
    # gather the owned rows in global numbering
    N = maps[0].size_global
    G = sp.lil_matrix((N, N))
    pattern = set()
    for A in mats:
        m0, m1 = A.index_maps
        for r in range(m0.size_local):
            gr = m0.local_range[0] + r
            gc = m1.local_to_global(A.cols[A.row_ptr[r]:A.row_ptr[r + 1]])
            G[gr, gc] = A.data[A.row_ptr[r]:A.row_ptr[r + 1]]
            pattern.update((gr, int(c)) for c in gc)
    # serial reference in the same global numbering
    x = M.box_vertices(nglob)
    cells = M.box_tets(nglob)
    # global index of each lattice vertex: owner's offset + owned lexicographic index
    gid = np.empty(len(x), dtype=np.int64)
    for i, m in zip(inputs, maps):
        part = i["part"]
        local, n_owned, ghosts, owners = part.vertex_numbering()
        ox, oy, oz = part.origin
        nz, ny, nx = local.shape
        izz, iyy, ixx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        lat = ((izz + oz) * (nglob[1] + 1) + (iyy + oy)) * (nglob[0] + 1) + (ixx + ox)
        lf = local.reshape(-1)
        own = lf < n_owned
        gid[lat.reshape(-1)[own]] = m.local_range[0] + lf[own]
    dm = gid[cells].astype(np.int32)
    ps = P.Problem(x, cells, dm, len(x), 1, "tetrahedron")
    pat, ref = P.oracle_assemble_matrix(O, ps, O.K_POISSON_P1_TET_A, constants=kappa)
    S = sp.csr_matrix((ref, pat.edges, pat.offsets), shape=(N, N))
    ref_pattern = set(zip(np.repeat(np.arange(N), np.diff(pat.offsets)).tolist(), pat.edges.tolist()))
    assert pattern == ref_pattern
    assert abs(G.tocsr() - S).max() <= 1e-12 * abs(S).max()
    assert O.matrix_squared_norm(mats) == pytest.approx(float(np.sum(ref**2)), rel=1e-12)
    # distributed mult == serial mult; x on the matrix column map
    maps1 = [A.index_maps[1] for A in mats]
    scs = O.make_scatterers(maps1, 1)
    xg = np.random.default_rng(0).random(N)
    xs = [np.concatenate([xg[m.local_range[0]:m.local_range[1]], np.zeros(m.num_ghosts)]) for m in maps1]
    ys = [np.zeros(A.index_maps[0].size_local + A.index_maps[0].num_ghosts) for A in mats]
    O.matrix_mult(mats, scs, xs, ys)
    yg = np.concatenate([y[: m.size_local] for y, m in zip(ys, maps)])
    assert np.max(np.abs(yg - S @ xg)) <= 1e-12 * np.max(np.abs(S @ xg))


def test_oracle_scatter_semantics(oracle):
    """python/test/unit/la/test_vector_scatter.py:26-113, common/test_scatterer.py:11-55:
    forward => ghosts == owner rank; reverse(add) => owners accumulate every ghost copy."""
    O = oracle
    inputs = brick_inputs(4, (4, 4, 2))
    maps = O.make_index_maps([i["n_owned"] for i in inputs], [i["ghosts"] for i in inputs], [i["owners"] for i in inputs])
    for bs in (1, 3):
        scs = O.make_scatterers(maps, bs)
        xs = [np.concatenate([np.full(bs * m.size_local, float(m.rank)), np.full(bs * m.num_ghosts, -1.0)]) for m in maps]
        O.vector_scatter_fwd(maps, scs, bs, xs)
        for m, x in zip(maps, xs):
            assert np.array_equal(x[bs * m.size_local:], np.repeat(m.owners.astype(float), bs))
        # reverse add: every ghost holds 1 -> owners gain the number of ranks ghosting each dof
        xs = [np.concatenate([np.zeros(bs * m.size_local), np.ones(bs * m.num_ghosts)]) for m in maps]
        O.vector_scatter_rev(maps, scs, bs, xs, "add")
        assert sum(x[: bs * m.size_local].sum() for m, x in zip(maps, xs)) == bs * sum(m.num_ghosts for m in maps)
        # plan invariants (common/Scatterer.h:98-101,154-187)
        for m, s in zip(maps, scs):
            assert np.all(np.diff(m.owners[s.remote_inds[::bs] // bs]) >= 0)
            assert s.displs_remote[-1] == bs * m.num_ghosts and len(s.local_inds) == s.displs_local[-1]


# ---------------------------------------------------------------------------------------------
# product host logic under gloo, world_size 2
# ---------------------------------------------------------------------------------------------
def _worker(rank, world, port, nglob, tmpdir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from dolfinx_b200 import common, la
        from oracle import oracle as O

        inputs = brick_inputs(world, nglob)
        me = inputs[rank]
        comm = common.Comm()
        assert (comm.rank, comm.size) == (rank, world)
        omaps, opats, omats = oracle_world(O, inputs)
        # IndexMap
        im = common.IndexMap(comm, me["n_owned"], me["ghosts"], me["owners"])
        om = omaps[rank]
        assert im.local_range == om.local_range and im.size_global == om.size_global
        assert np.array_equal(im.src, om.src) and np.array_equal(im.dest, om.dest)
        # Scatterer, bs 1 and 3
        for bs in (1, 3):
            sc = common.Scatterer(im, bs)
            osc = O.make_scatterers(omaps, bs)[rank]
            for name in ("local_inds", "remote_inds", "sizes_local", "displs_local", "sizes_remote", "displs_remote"):
                assert np.array_equal(getattr(sc, name), getattr(osc, name)), (name, bs)
        # SparsityPattern (host COO path) — bit-exact graph, off-diagonal offsets and extended column map
        pat = la.SparsityPattern(comm, [im, im], [1, 1])
        pat.insert_cells(np.arange(len(me["dofmap"])), me["dofmap"], me["dofmap"])
        pat.finalize()
        op = opats[rank]
        edges, offsets = pat.graph
        assert np.array_equal(offsets, op.offsets) and np.array_equal(edges, op.edges)
        assert np.array_equal(pat.off_diagonal_offsets, op.off_diagonal_offsets)
        m1, om1 = pat.index_map(1), op.index_maps[1]
        assert np.array_equal(m1.ghosts, om1.ghosts) and np.array_equal(m1.owners, om1.owners)
        assert np.array_equal(m1.src, om1.src) and np.array_equal(m1.dest, om1.dest)
        # MatrixCSR ghost-row plan
        plan = la.matrix_ghost_plan([pat.index_map(0), pat.index_map(1)], (1, 1), offsets, edges)
        oA = omats[rank]
        assert np.array_equal(plan["ghost_row_to_rank"], oA.ghost_row_to_rank)
        assert np.array_equal(plan["val_send_disp"], oA.val_send_disp)
        assert np.array_equal(plan["val_recv_disp"], oA.val_recv_disp)
        assert np.array_equal(plan["unpack_pos"], oA.unpack_pos)
        # la::transpose (la/mattrans.h): the host side (exchange of the ghost-column entries, merged structure, column
        # map of the result) against the oracle's simulated ranks; the local transpose (a device kernel in the product)
        # is taken from the oracle here
        for tbs in ((1, 1), (2, 3)):
            nb = tbs[0] * tbs[1]
            tmats = []
            for r_, A_ in enumerate(omats):
                d_ = np.random.default_rng(100 + r_).random(len(A_.cols) * nb)
                tmats.append(O.OMatrix(A_.index_maps, tbs, d_, A_.cols, A_.row_ptr, A_.off_diag_offset))
            oT = O.transpose(tmats)[rank]
            mine = tmats[rank]
            c0, rp0, v0 = O.local_transpose(mine)
            tp = la.matrix_transpose_plan(pat.index_map(0), pat.index_map(1), tbs, mine.row_ptr, mine.cols,
                                          mine.off_diag_offset, rp0, c0,
                                          lambda ks: mine.data.reshape(-1, nb)[ks].reshape(-1))
            assert np.array_equal(tp["row_ptr"], oT.row_ptr) and np.array_equal(tp["cols"], oT.cols)
            assert np.array_equal(tp["row_ptr"][:-1] + tp["off_diag"], oT.off_diag_offset)
            tr, tc = tp["maps"]
            assert tr.num_ghosts == 0 and tr.local_range == oT.index_maps[0].local_range
            assert np.array_equal(tc.ghosts, oT.index_maps[1].ghosts) and np.array_equal(tc.owners, oT.index_maps[1].owners)
            vals = np.zeros(len(tp["cols"]) * nb).reshape(-1, nb)
            vals[tp["local_dst"]] = v0.reshape(-1, nb)
            if len(tp["recv_dst"]):
                vals[tp["recv_dst"]] = tp["recv_blocks"]
            assert np.array_equal(vals.reshape(-1), oT.data)
        # la::matmul (la/matmul.h): fetch_ghost_rows on the host (rows of B behind the ghost columns of A, the
        # column map of C) against the oracle's simulated ranks, then impl::matmul on the fetched rows: C = A A
        mmats = []
        for r_, A_ in enumerate(omats):
            d_ = np.random.default_rng(200 + r_).random(len(A_.cols))
            # square product: rows of B = owned columns of A -> drop the ghost rows of the assembled structure
            nl = A_.index_maps[0].size_local
            n = int(A_.row_ptr[nl])
            rowmap = O.make_index_maps([m.index_maps[0].size_local for m in omats], [[] for _ in omats], [[] for _ in omats])[r_]
            mmats.append(O.OMatrix([rowmap, A_.index_maps[1]], (1, 1), d_[:n], A_.cols[:n], A_.row_ptr[: nl + 1],
                                   A_.off_diag_offset[:nl]))
        oC, oghost = O.matmul(mmats, mmats, return_ghost_rows=True)
        mine = mmats[rank]
        rmap = common.IndexMap(comm, mine.index_maps[0].size_local)
        mp = la.matrix_matmul_plan(pat.index_map(1), rmap, pat.index_map(1), mine.row_ptr, mine.cols,
                                   lambda ks: mine.data[ks])
        ocm = oC[rank].index_maps[1]
        assert np.array_equal(mp["col_map"].ghosts, ocm.ghosts) and np.array_equal(mp["col_map"].owners, ocm.owners)
        assert np.array_equal(mp["ghost_row_ptr"], oghost[rank][0]) and np.array_equal(mp["ghost_cols"], oghost[rank][1])
        assert np.array_equal(mp["ghost_vals"], oghost[rank][2])
        rp_, od_, cols_, vals_ = O.matmul_local(mine, mine, ocm, mp["ghost_row_ptr"], mp["ghost_cols"], mp["ghost_vals"])
        assert np.array_equal(rp_, oC[rank].row_ptr) and np.array_equal(cols_, oC[rank].cols)
        assert np.array_equal(vals_, oC[rank].data)
        # collectives used by la::norm / inner_product
        assert comm.allreduce_sum(float(rank + 1)) == sum(range(1, world + 1))
        assert comm.allreduce_max(float(rank)) == world - 1
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nglob", [(2, (4, 3, 2)), (3, (6, 2, 2))])
def test_product_host_plans_gloo(tmp_path, world, nglob):
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, nglob, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


# ---------------------------------------------------------------------------------------------
# vectorised host helpers against literal per-entry loops (the loops of the reference)
# ---------------------------------------------------------------------------------------------
def test_host_plan_helpers_against_loops():
    from dolfinx_b200 import la

    rng = np.random.default_rng(7)
    for trial in range(20):
        # first_occurrences_by_row: insertion-ordered distinct columns per row (SparsityPattern.cpp:291-330)
        nrows, ncols, n = int(rng.integers(1, 9)), int(rng.integers(1, 12)), int(rng.integers(0, 200))
        rows, cols = rng.integers(0, nrows, n), rng.integers(0, ncols, n)
        lists = [[] for _ in range(nrows)]
        for r, c in zip(rows, cols):
            if c not in lists[r]:
                lists[r].append(int(c))
        r2, c2 = la.first_occurrences_by_row(rows, cols, ncols)
        assert np.array_equal(r2, np.repeat(np.arange(nrows), [len(x) for x in lists]))
        assert np.array_equal(c2, np.array([c for x in lists for c in x], dtype=np.int64))

        # received_columns_to_local: new ghost columns in arrival order (SparsityPattern.cpp:389-423)
        lo = int(rng.integers(0, 50))
        hi = lo + int(rng.integers(1, 20))
        known = rng.choice(np.setdiff1d(np.arange(0, 120), np.arange(lo, hi)), size=int(rng.integers(0, 10)), replace=False)
        owners = rng.integers(0, 4, known.size)
        m = int(rng.integers(0, 150))
        in_cols = rng.integers(0, 120, m)
        in_own = in_cols % 5  # the owner is a function of the column, as in a real exchange
        g2l = {int(g): (hi - lo) + i for i, g in enumerate(known)}
        gh, go, out = list(known), list(owners), []
        for c, o in zip(in_cols, in_own):
            if lo <= c < hi:
                out.append(int(c) - lo)
            else:
                if int(c) not in g2l:
                    g2l[int(c)] = (hi - lo) + len(gh)
                    gh.append(int(c))
                    go.append(int(o))
                out.append(g2l[int(c)])
        o2, gh2, go2 = la.received_columns_to_local(in_cols, in_own, (lo, hi), list(known), list(owners))
        assert np.array_equal(o2, np.array(out, dtype=np.int32)) and list(gh2) == gh and list(go2) == go

        # locate_entries: lower_bound per entry (MatrixCSR.h:829-845)
        nr, nc = int(rng.integers(1, 15)), int(rng.integers(1, 20))
        indptr, indices = [0], []
        for _ in range(nr):
            cs = np.sort(rng.choice(nc, size=int(rng.integers(0, nc + 1)), replace=False))
            indices += cs.tolist()
            indptr.append(len(indices))
        indptr, indices = np.array(indptr, dtype=np.int64), np.array(indices, dtype=np.int32)
        if len(indices):
            k = rng.integers(0, len(indices), int(rng.integers(0, 60)))
            qr = np.searchsorted(indptr, k, side="right") - 1
            got = la.locate_entries(indptr, indices, qr, indices[k], nc)
            assert np.array_equal(got, k)
            # an entry that is not in the pattern is reported
            full = [(r, c) for r in range(nr) for c in range(nc) if c not in indices[indptr[r]:indptr[r + 1]]]
            if full:
                r, c = full[int(rng.integers(0, len(full)))]
                assert la.locate_entries(indptr, indices, np.append(qr, r), np.append(indices[k], c), nc) is None
        assert la.locate_entries(indptr, indices, [], [], nc).size == 0


@pytest.mark.parametrize("nranks,nglob", [(2, (4, 3, 2)), (4, (4, 4, 2)), (8, (4, 4, 4)), (3, (5, 2, 2))])
def test_oracle_distributed_p2_equals_serial(oracle, nranks, nglob):
    """BASELINE configs[2] on several ranks: the P2 brick partition (mesh.p2_partition: the dofs of a Kuhn box are the
    points of the doubled lattice) assembled per simulated rank + scatter_rev equals the serial P2 assembly entry by
    entry, the dofs matched through their coordinates (cf. cpp/test/matrix.cpp:59-64)."""
    O = oracle
    pg = M.pgrid_for(nranks)
    parts = [M.BoxPartition(r, pg, nglob) for r in range(nranks)]
    inputs = []
    for part in parts:
        xd, dm, n_owned, ghosts, owners, dc = M.p2_partition(part)
        inputs.append(dict(x=part.coordinates(), x_dofmap=xd, dofmap=dm, n_owned=n_owned, ghosts=ghosts, owners=owners, dc=dc))
    maps, pats, mats = oracle_world(O, inputs)
    kappa = np.array([2.0])
    for i, A in zip(inputs, mats):
        O.assemble_matrix(O.K_POISSON_P2_TET_A, i["x_dofmap"], i["x"], np.arange(len(i["dofmap"])), i["dofmap"], 1,
                          i["dofmap"], 1, A.data, A.cols, A.row_ptr, constants=kappa)
    O.matrix_scatter_rev(mats)
    N = maps[0].size_global
    assert N == (2 * nglob[0] + 1) * (2 * nglob[1] + 1) * (2 * nglob[2] + 1)
    # doubled-lattice key of every global dof, from the owners' coordinates
    n2 = 2 * np.array(nglob)
    key_of_global = np.empty(N, dtype=np.int64)
    for i, m in zip(inputs, maps):
        k = np.rint(i["dc"][: m.size_local] * n2).astype(np.int64)
        key_of_global[m.local_range[0]:m.local_range[1]] = (k[:, 2] * (n2[1] + 1) + k[:, 1]) * (n2[0] + 1) + k[:, 0]
    assert len(np.unique(key_of_global)) == N
    for i, m in zip(inputs, maps):  # ghost global indices point at dofs with the same coordinates
        kg = np.rint(i["dc"][m.size_local:] * n2).astype(np.int64)
        assert np.array_equal(key_of_global[m.ghosts], (kg[:, 2] * (n2[1] + 1) + kg[:, 1]) * (n2[0] + 1) + kg[:, 0])
    G = sp.lil_matrix((N, N))
    for A in mats:
        m0, m1 = A.index_maps
        for r in range(m0.size_local):
            gc = m1.local_to_global(A.cols[A.row_ptr[r]:A.row_ptr[r + 1]])
            G[m0.local_range[0] + r, gc] = A.data[A.row_ptr[r]:A.row_ptr[r + 1]]
    # serial reference, renumbered to the distributed global numbering through the coordinates
    ps = P.tet_p2(nglob)
    ks = np.rint(ps.dof_coords * n2).astype(np.int64)
    key_s = (ks[:, 2] * (n2[1] + 1) + ks[:, 1]) * (n2[0] + 1) + ks[:, 0]
    glob_of_key = np.empty(N, dtype=np.int64)
    glob_of_key[key_of_global] = np.arange(N)
    to_glob = glob_of_key[key_s]  # serial dof -> distributed global dof
    pat, ref = P.oracle_assemble_matrix(O, ps, O.K_POISSON_P2_TET_A, constants=kappa)
    S = sp.csr_matrix((ref, pat.edges, pat.offsets), shape=(N, N)).tocoo()
    S = sp.csr_matrix((S.data, (to_glob[S.row], to_glob[S.col])), shape=(N, N))
    Gc = G.tocsr()
    assert abs(Gc - S).max() <= 1e-12 * abs(S).max()
    # same structure: the index sets of the owned rows agree with the serial pattern (explicit zeros included)
    keys_d = []
    for A in mats:
        m0, m1 = A.index_maps
        nl = m0.size_local
        rows_g = np.repeat(np.arange(nl, dtype=np.int64) + m0.local_range[0], np.diff(A.row_ptr[: nl + 1]))
        keys_d.append(rows_g * N + m1.local_to_global(A.cols[: A.row_ptr[nl]]))
    rows_s = np.repeat(np.arange(N, dtype=np.int64), np.diff(pat.offsets))
    keys_s = to_glob[rows_s] * N + to_glob[pat.edges]
    assert np.array_equal(np.sort(np.concatenate(keys_d)), np.sort(keys_s))
    assert O.matrix_squared_norm(mats) == pytest.approx(float(np.sum(ref**2)), rel=1e-12)


# ---------------------------------------------------------------------------------------------
# the reference's ring fixtures: cpp/test/vector.cpp:22-55, cpp/test/common/index_map.cpp:22-170
# ---------------------------------------------------------------------------------------------
def ring_maps(O, size, size_local=100):
    """create_index_map of cpp/test/common/index_map.cpp:22-38: (size - 1) * 3 ghosts owned by the next rank."""
    ghosts = [[((r + 1) % size) * size_local + i for i in range((size - 1) * 3)] for r in range(size)]
    owners = [[(r + 1) % size] * len(g) for r, g in enumerate(ghosts)]
    return O.make_index_maps([size_local] * size, ghosts, owners)


@pytest.mark.parametrize("size", [1, 2, 3, 5])
def test_oracle_ring_vector_and_scatter(oracle, size):
    O = oracle
    size_local = 100
    maps = ring_maps(O, size, size_local)
    # cpp/test/vector.cpp:40-54: closed-form norms (ghosts do not count)
    ones = [np.ones(m.size_local + m.num_ghosts) for m in maps]
    assert O.inner_product(maps, 1, ones, ones) == size * size_local
    vs = [np.full(m.size_local + m.num_ghosts, float(m.rank)) for m in maps]
    sumn2 = size_local * (size - 1) * size * (2 * size - 1) / 6
    assert O.inner_product(maps, 1, vs, vs) == sumn2
    for n in (1, 5, 10):
        scs = O.make_scatterers(maps, n)
        # index_map.cpp:40-100: forward scatter of val * rank => every ghost holds val * next rank
        val = 11.0
        xs = [np.concatenate([np.full(n * m.size_local, val * m.rank), np.full(n * m.num_ghosts, -1.0)]) for m in maps]
        O.vector_scatter_fwd(maps, scs, n, xs)
        for m, x in zip(maps, xs):
            assert np.all(x[n * m.size_local:] == val * ((m.rank + 1) % size))
        # index_map.cpp:103-166: reverse scatter (add) of `value` in every ghost => sum = n * value * num_ghosts, twice
        value = 15.0
        xs = [np.concatenate([np.zeros(n * m.size_local), np.full(n * m.num_ghosts, value)]) for m in maps]
        for rep in (1, 2):
            O.vector_scatter_rev(maps, scs, n, xs, "add")
            for m, x in zip(maps, xs):
                # rank r receives from the previous rank's ghosts
                assert x[: n * m.size_local].sum() == rep * n * value * maps[(m.rank - 1) % size].num_ghosts


def _ring_worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from dolfinx_b200 import common
        from oracle import oracle as O

        omaps = ring_maps(O, world)
        om = omaps[rank]
        im = common.IndexMap(common.Comm(), om.size_local, om.ghosts, om.owners)
        assert im.local_range == om.local_range and im.size_global == world * 100
        assert np.array_equal(im.src, om.src) and np.array_equal(im.dest, om.dest)
        assert np.array_equal(im.local_to_global(np.arange(om.size_local + om.num_ghosts)),
                              om.local_to_global(np.arange(om.size_local + om.num_ghosts)))
        for n in (1, 5, 10):
            sc = common.Scatterer(im, n)
            osc = O.make_scatterers(omaps, n)[rank]
            for name in ("local_inds", "remote_inds", "sizes_local", "displs_local", "sizes_remote", "displs_remote"):
                assert np.array_equal(getattr(sc, name), getattr(osc, name)), (name, n)
            assert len(sc.local_indices()) == n * 3 * (world - 1) and len(sc.remote_indices()) == n * om.num_ghosts
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_product_ring_index_map_and_scatterer_gloo(tmp_path, world):
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_ring_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
