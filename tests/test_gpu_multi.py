"""Multi-GPU parity (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

One process per GPU (torch.distributed, NCCL).  Every rank builds its brick of the box, runs the
product path (device sparsity build with the ghost-row exchange, assembly, MatrixCSR::scatter_rev,
Vector scatter, distributed mult through NCCL send/recv) and compares with the oracle's simulated
ranks: structure and plans bit-exact, values to 1e-12."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, nglob, cell, tmpdir):
    import faulthandler

    faulthandler.dump_traceback_later(150, exit=True)  # a hung collective must not eat the GPU budget
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from dolfinx_b200 import _lib as K
        from dolfinx_b200 import common, fem, la
        from oracle import oracle as O
        from tests.test_distributed_host import brick_inputs, oracle_world

        bs = 1 if cell == "tet" else 3
        inputs = brick_inputs(world, nglob, cell)
        me = inputs[rank]
        comm = common.Comm()
        omaps, opats, omats = oracle_world(O, inputs, bs)
        if cell == "tet":
            kA, okA, kL, okL, consts = K.K_POISSON_P1_TET_A, O.K_POISSON_P1_TET_A, K.K_LOAD_P1_TET_L, O.K_LOAD_P1_TET_L, [2.0]
        else:
            kA, okA, kL, okL, consts = K.K_ELASTICITY_Q1_HEX_A, O.K_ELASTICITY_Q1_HEX_A, K.K_LOAD_Q1_HEX_L, O.K_LOAD_Q1_HEX_L, [[1.0, 1.5]]
        cc = np.asarray(consts, dtype=np.float64).reshape(-1)

        im = common.IndexMap(comm, me["n_owned"], me["ghosts"], me["owners"])
        msh = fem.Mesh(comm, me["x"], me["x_dofmap"], cell)
        V = fem.FunctionSpace(msh, "Lagrange", fem.DofMap(me["dofmap"], bs, im))
        a = fem.Form([V, V], {fem.IntegralType.cell: [(0, kA, None, [])]}, constants=[fem.Constant(c) for c in consts])
        sp = fem.create_sparsity_pattern(a)  # device cell list -> bfx_sparsity_ghost_rows + bfx_sparsity_build
        sp.finalize()
        op = opats[rank]
        edges, offsets = sp.graph
        assert np.array_equal(offsets, op.offsets) and np.array_equal(edges, op.edges)
        assert np.array_equal(sp.off_diagonal_offsets, op.off_diagonal_offsets)
        assert np.array_equal(sp.index_map(1).ghosts, op.index_maps[1].ghosts)
        assert np.array_equal(sp.index_map(1).owners, op.index_maps[1].owners)

        A = la.MatrixCSR(sp)
        oA = omats[rank]
        for name in ("ghost_row_to_rank", "val_send_disp", "val_recv_disp", "unpack_pos"):
            assert np.array_equal(A._plan_arrays[name], getattr(oA, name)), name
        fem.assemble_matrix(A, a)
        for i, M in zip(inputs, omats):
            O.assemble_matrix(okA, i["x_dofmap"], i["x"], np.arange(len(i["dofmap"])), i["dofmap"], bs, i["dofmap"], bs,
                              M.data, M.cols, M.row_ptr, constants=cc)
        pre = A.data.cpu().numpy()
        scale = np.max(np.abs(oA.data))
        assert np.max(np.abs(pre - oA.data)) <= 1e-12 * scale
        A.scatter_reverse()
        O.matrix_scatter_rev(omats)
        post = A.data.cpu().numpy()
        assert np.max(np.abs(post - oA.data)) <= 1e-12 * scale
        assert np.all(post[A.indptr[im.size_local] * bs * bs:] == 0)
        assert A.squared_norm() == pytest.approx(O.matrix_squared_norm(omats), rel=1e-12)
        # overlapped variant: boundary cells -> start ghost-row exchange -> interior cells -> add
        A2 = la.MatrixCSR(sp)
        fem.assemble_matrix_overlapped(A2, a)
        assert np.max(np.abs(A2.data.cpu().numpy() - oA.data)) <= 1e-12 * scale
        bnd, interior = fem._boundary_interior_cells(a, a.integral(fem.IntegralType.cell, 0))
        assert interior.numel() > 0 and bnd.numel() + interior.numel() == len(me["dofmap"])
        # only ranks with ghost rows have boundary cells (the lowest brick owns its whole shared plane)
        assert (bnd.numel() > 0) == (im.num_ghosts > 0)

        # vector assembly + reverse scatter (la/Vector.h:371-379)
        f = fem.Function(V)
        fh = [np.random.default_rng(100 + r).random(bs * (m.size_local + m.num_ghosts)) for r, m in enumerate(omaps)]
        oscs = O.make_scatterers(omaps, bs)
        O.vector_scatter_fwd(omaps, oscs, bs, fh)  # consistent ghost values
        f.x.array.copy_(torch.from_numpy(fh[rank]))
        f.x.array[bs * im.size_local:] = -1.0
        f.x.scatter_forward()
        assert np.array_equal(f.x.array.cpu().numpy(), fh[rank])
        L = fem.Form([V], {fem.IntegralType.cell: [(0, kL, None, [0])]}, coefficients=[f])
        b = la.Vector(im, bs)
        fem.assemble_vector(b, L)
        b.scatter_reverse(la.InsertMode.add)
        obs = []
        for i, m, fr in zip(inputs, omaps, fh):
            ob = np.zeros(bs * (m.size_local + m.num_ghosts))
            cells = np.arange(len(i["dofmap"]))
            coeffs = np.zeros((len(cells), i["dofmap"].shape[1] * bs))
            O.pack_coefficient(coeffs, 0, fr, i["dofmap"], bs, cells=cells)
            O.assemble_vector(okL, i["x_dofmap"], i["x"], cells, i["dofmap"], bs, ob, coeffs=coeffs)
            obs.append(ob)
        O.vector_scatter_rev(omaps, oscs, bs, obs, "add")
        n0 = bs * im.size_local
        assert np.max(np.abs(b.array.cpu().numpy()[:n0] - obs[rank][:n0])) <= 1e-12 * np.max(np.abs(obs[rank]))
        assert la.norm(b) == pytest.approx(np.sqrt(O.inner_product(omaps, bs, obs, obs)), rel=1e-12)

        # distributed mult on the matrix column map, overlapped with the forward scatter
        maps1 = [M.index_maps[1] for M in omats]
        scs1 = O.make_scatterers(maps1, bs)
        xs = [np.concatenate([np.random.default_rng(7 + r).random(bs * m.size_local), np.zeros(bs * m.num_ghosts)])
              for r, m in enumerate(maps1)]
        ys = [np.zeros(bs * (M.index_maps[0].size_local + M.index_maps[0].num_ghosts)) for M in omats]
        x = la.Vector(A.index_map(1), bs)
        y = la.Vector(A.index_map(0), bs)
        x.array.copy_(torch.from_numpy(xs[rank]))
        A.mult(x, y)
        O.matrix_mult(omats, scs1, xs, ys)
        assert np.max(np.abs(y.array.cpu().numpy()[:n0] - ys[rank][:n0])) <= 1e-12 * np.max(np.abs(ys[rank]))
        torch.cuda.synchronize()
        dist.barrier()
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    except BaseException:
        # a failed rank must not leave its peers waiting inside a collective: report and leave at once
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
    dist.destroy_process_group()


@pytest.mark.parametrize("cell,nglob", [("tet", (8, 6, 5)), ("hex", (6, 5, 4))])
def test_multi_gpu_parity(tmp_path, cell, nglob):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, nglob, cell, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
