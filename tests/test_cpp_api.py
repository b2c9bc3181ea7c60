"""The header-only C++ mirror (dolfinx_b200/cpp/dolfinx_b200.h) compiles against include/bfx.h with
g++ -std=c++20 -Wall -Wextra -Werror and — on a GPU — reproduces the reference's demo/test flow."""

import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "test_cpp_api")
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "test_cpp_api.cpp"),
           "-o", exe, "-L", os.path.join(ROOT, "dolfinx_b200"), "-lbfx", "-Wl,-rpath," + os.path.join(ROOT, "dolfinx_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    # without a GPU the program must stop at the device probe (exit 77), never compute on the CPU
    assert r.returncode in (0, 77), r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CPP_API_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_multirank_on_gpus(tmp_path, oracle):
    """The same program with one GPU per simulated rank: NCCL communicator per thread, la::Vector scatter_fwd /
    scatter_rev and MatrixCSR::scatter_rev / squared_norm of the C++ mirror on the plans it built (2 GPUs needed)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    test_cpp_mirror_multirank_host_plans(tmp_path, oracle, "hex", 2, (4, 3, 3), 3, device=True)


@pytest.mark.parametrize("cell,ranks,nglob,bs", [("tet", 3, (5, 4, 3), 1), ("hex", 4, (4, 4, 3), 3), ("tet", 2, (3, 2, 2), 2)])
def test_cpp_mirror_multirank_host_plans(tmp_path, oracle, cell, ranks, nglob, bs, device=False):
    """The C++ mirror on several ranks (SURVEY.md §8 rows a11-a15 on the host side): IndexMap, Scatterer plan,
    SparsityPattern::finalize and the MatrixCSR ghost-row plan computed by dolfinx_b200.h on simulated ranks (threads
    with in-memory exchange callbacks) are bit-identical to the oracle's restatement of the reference constructors."""
    import numpy as np

    from dolfinx_b200 import mesh as M

    exe = str(tmp_path / "test_cpp_multirank")
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread",
           os.path.join(ROOT, "tests", "cpp", "test_cpp_multirank.cpp"), "-o", exe, "-L", os.path.join(ROOT, "dolfinx_b200"),
           "-lbfx", "-Wl,-rpath," + os.path.join(ROOT, "dolfinx_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pg = M.pgrid_for(ranks)
    ins = []
    for rank in range(ranks):
        part = M.BoxPartition(rank, pg, nglob)
        xd, dm, n_owned, ghosts, owners = part.tets() if cell == "tet" else part.hexes()
        rows, cols = oracle.sparsity_insert_cells(np.arange(len(dm)), dm, dm)
        ins.append((n_owned, ghosts, owners, rows, cols))
    with open(tmp_path / "in.txt", "w") as f:
        f.write(f"{ranks} {bs}\n")
        for n_owned, ghosts, owners, rows, cols in ins:
            f.write(f"{n_owned} {len(ghosts)}\n" + " ".join(map(str, ghosts)) + "\n" + " ".join(map(str, owners)) + "\n")
            f.write(f"{len(rows)}\n" + " ".join(map(str, rows)) + "\n" + " ".join(map(str, cols)) + "\n")
    r = subprocess.run([exe, str(tmp_path / "in.txt"), str(tmp_path / "out.txt")] + (["device"] if device else []),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CPP_MULTIRANK_OK" in r.stdout, r.stdout + r.stderr
    assert not device or "device path on" in r.stdout
    got, cur = [], None
    for line in open(tmp_path / "out.txt"):
        t = line.split()
        if t[0] == "rank":
            cur = {}
            got.append(cur)
        elif t[0] == "range":
            cur["range"] = [int(v) for v in t[1:]]
        else:
            assert int(t[1]) == len(t) - 2
            cur[t[0]] = np.array([int(v) for v in t[2:]], dtype=np.int64)
    # the oracle on the same inputs
    maps = oracle.make_index_maps([i[0] for i in ins], [i[1] for i in ins], [i[2] for i in ins])
    scs = oracle.make_scatterers(maps, bs)
    pats = oracle.sparsity_finalize(maps, maps, (bs, bs), [i[3] for i in ins], [i[4] for i in ins])
    mats = oracle.make_matrices(pats)
    for rank in range(ranks):
        g, m, sc, p, A = got[rank], maps[rank], scs[rank], pats[rank], mats[rank]
        assert g["range"] == [m.local_range[0], m.local_range[1], m.size_global]
        for name, ref in [("src", m.src), ("dest", m.dest), ("local_inds", sc.local_inds), ("remote_inds", sc.remote_inds),
                          ("sizes_local", sc.sizes_local), ("displs_local", sc.displs_local),
                          ("sizes_remote", sc.sizes_remote), ("displs_remote", sc.displs_remote), ("edges", p.edges),
                          ("offsets", p.offsets), ("off_diag", p.off_diagonal_offsets),
                          ("col_ghosts", p.index_maps[1].ghosts), ("col_owners", p.index_maps[1].owners),
                          ("col_src", p.index_maps[1].src), ("col_dest", p.index_maps[1].dest),
                          ("ghost_row_to_rank", A.ghost_row_to_rank), ("val_send_disp", A.val_send_disp),
                          ("val_recv_disp", A.val_recv_disp), ("unpack_pos", A.unpack_pos)]:
            assert np.array_equal(g[name], np.asarray(ref, dtype=np.int64)), (rank, name)
