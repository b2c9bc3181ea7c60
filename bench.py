#!/usr/bin/env python
"""Benchmark of the DOLFINx assembly hot path on B200 (contract: task statement, section 4).

    python bench.py --gpus N --steps K --warmup W           # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

A *step* is one full matrix assembly call sequence on one batch of synthetic input
(SURVEY.md §8d), as a time loop issues it: ``A.set_value(0)`` + ``assemble_matrix`` with Dirichlet
markers + ``set_diagonal`` + ``MatrixCSR.scatter_reverse`` (overlapped with the interior cells on
N > 1 GPUs).  ``value`` = global DOFs / step time, inputs resident in HBM.  ``roofline`` times the
assembly kernel alone.  ``e2e`` is the same metric through the host-buffer C-ABI entry
(bfx_assemble_matrix_cells_host_begin/_end, two steps in flight): geometry and bc markers copied
host->device and all CSR values copied device->host inside the timed region.
SpMV (MatrixCSR::mult) is timed in a second loop and reported under ``spmv``.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "assembled_dofs_per_s_fp64_matrix"
UNIT = "DOF/s"

CONFIGS = {
    # name: (cell, element, default n, kernel ids (A, L), block size)
    "p1": ("tet", "P1", 256),
    "p2": ("tet", "P2", 128),
    "q1": ("hex", "Q1", 192),
}


def measured_traffic(config, strategy, n):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f).get(f"{config}:{strategy}:{n}")
        return None if t is None else t["bytes"]
    except (OSError, ValueError, KeyError):
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# problem construction (inputs of the path; not timed)
# ---------------------------------------------------------------------------------------------
def build_problem(cfg, n, comm, device):
    """Returns dict with mesh, V, forms, bc, counts.  Single rank: whole box; N ranks: brick partition."""
    import torch

    from dolfinx_b200 import _lib as K
    from dolfinx_b200 import common, fem, mesh as M

    cell, elem, _ = CONFIGS[cfg]
    if comm.size == 1:
        n3 = (n, n, n) if np.isscalar(n) else tuple(n)
        x = M.box_vertices_torch(n3, device)
        if cell == "tet":
            cells = M.box_tets_torch(n3, device)
        else:
            cells = M.box_hexes_torch(n3, device)
        nv = x.shape[0]
        if elem in ("P1", "Q1"):
            new = M.first_touch_numbering_torch(cells, nv)
            dofmap = new[cells.long()].contiguous()
            ndofs = nv
            dof_x = torch.empty_like(x)
            dof_x[new.long()] = x
            del new
        else:  # P2: vertices + edges, numbered on the host helper then first-touch
            cells_h = cells.cpu().numpy()
            dm_h, ndofs = M.p2_tet_dofmap(cells_h, nv)
            dofmap = torch.from_numpy(dm_h).to(device)
            # dof nodes: vertices, then edge midpoints (Basix edge order)
            dof_x = torch.empty((ndofs, 3), dtype=torch.float64, device=device)
            cl = cells.long()
            dof_x[dofmap[:, :4].reshape(-1).long()] = x[cl.reshape(-1)]
            for k, (va, vb) in enumerate(M.TET_EDGES):
                dof_x[dofmap[:, 4 + k].long()] = 0.5 * (x[cl[:, int(va)]] + x[cl[:, int(vb)]])
            del cl
        im = common.IndexMap(comm, ndofs)
        n_cells_local = cells.shape[0]
        msh = fem.Mesh(comm, x, cells, cell, n_cells_local)
        ghosts = None
    else:
        pg = M.pgrid_for(comm.size)
        nglob = tuple(n * pg[d] for d in range(3))
        part = M.BoxPartition(comm.rank, pg, nglob)
        if elem == "P2":
            # BASELINE configs[2] on N GPUs: the P2 dofs of a Kuhn box are the points of the doubled lattice, partitioned
            # like vertices (mesh.p2_partition; host numpy, then uploaded)
            xd_h, dm_h, n_owned, ghosts, owners, dc_h = M.p2_partition(part)
            x = torch.from_numpy(part.coordinates()).to(device)
            cells = torch.from_numpy(xd_h).to(device)
            dofmap = torch.from_numpy(dm_h).to(device)
            dof_x = torch.from_numpy(dc_h).to(device)
            ndofs = n_owned + len(ghosts)
            del xd_h, dm_h, dc_h
        else:
            x, cells, dofmap, local, n_owned, ghosts, owners = M.partition_cells_torch(part, cell, device)
            ndofs = n_owned + len(ghosts)
            dof_x = torch.empty((ndofs, 3), dtype=torch.float64, device=device)
            dof_x[local.reshape(-1).long()] = x
            del local
        im = common.IndexMap(comm, n_owned, ghosts, owners)
        msh = fem.Mesh(comm, x, cells, cell, cells.shape[0])
    bs = 3 if elem == "Q1" else 1
    V = fem.FunctionSpace(msh, elem, fem.DofMap(dofmap, bs, im))
    if elem == "P1":
        kA, kL, consts = K.K_POISSON_P1_TET_A, K.K_LOAD_P1_TET_L, [2.0]
    elif elem == "P2":
        kA, kL, consts = K.K_POISSON_P2_TET_A, K.K_LOAD_P2_TET_L, [2.0]
    else:
        E, nu = 1.0e9, 0.3
        kA, kL = K.K_ELASTICITY_Q1_HEX_A, K.K_LOAD_Q1_HEX_L
        consts = [[E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))]]
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, kA, None, [])]}, constants=[fem.Constant(c) for c in consts])
    f = fem.Function(V)
    if dof_x is not None:
        if bs == 1:
            f.x.array.copy_(10.0 * torch.exp(-((dof_x[:, 0] - 0.5) ** 2 + (dof_x[:, 1] - 0.5) ** 2) / 0.02))
        else:
            fv = f.x.array.view(-1, 3)
            fv[:, 0] = 10 * 300.0**2 * dof_x[:, 0]
            fv[:, 1] = 10 * 300.0**2 * dof_x[:, 1]
    else:
        f.x.array.fill_(1.0)
    L = fem.Form([V], {fem.IntegralType.cell: [(0, kL, None, [0])]}, coefficients=[f])
    # Dirichlet dofs: x0 in {0, 1} (cpp/demo/poisson/main.cpp:159-175); elasticity: x0 = 0 or x1 = 1
    if dof_x is not None:
        if elem == "Q1":
            mask = (dof_x[:, 0] < 1e-12) | (dof_x[:, 1] > 1 - 1e-12)
        else:
            mask = (dof_x[:, 0] < 1e-12) | (dof_x[:, 0] > 1 - 1e-12)
        bdofs = torch.nonzero(mask).reshape(-1).to(torch.int32).cpu().numpy()
    else:
        bdofs = np.zeros(0, dtype=np.int32)
    # (the matrix does not see g; apply_lifting does: g = 1, 2, .. per component)
    bc = fem.DirichletBC(fem.Constant(1.0 + np.arange(bs)), bdofs, V)
    return dict(mesh=msh, V=V, a=a, L=L, bc=bc, ndofs_local=im.size_local, ndofs_global=im.size_global,
                n_cells=int(cells.shape[0]), n_x=int(x.shape[0]), bs=bs, nx=int(cells.shape[1]), nd=int(dofmap.shape[1]))


def alg_bytes_asm(pb, nnz, n_rows):
    """B_asm of SURVEY.md §8d."""
    bs = pb["bs"]
    return (4 * pb["n_cells"] * (pb["nx"] + pb["nd"]) + 24 * pb["n_x"] + 8 * bs * bs * nnz + 4 * nnz
            + 8 * (n_rows + 1) + bs * n_rows)


def alg_bytes_spmv(pb, nnz, n_rows, n_cols):
    bs = pb["bs"]
    return (8 * bs * bs + 4) * nnz + 8 * (n_rows + 1) + 8 * bs * n_cols + 16 * bs * n_rows


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores, decomposed like MPI ranks
# ---------------------------------------------------------------------------------------------
def _mesh_fixture():
    """The numpy box / brick-partition generators (dolfinx_b200/mesh.py), loaded as a stand-alone file: the CPU arm
    imports neither the dolfinx_b200 package nor its native library."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bfx_mesh_fixture", os.path.join(ROOT, "dolfinx_b200", "mesh.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bfx_mesh_fixture"] = mod
    spec.loader.exec_module(mod)
    return mod


def _grid3(n):
    """n = px * py * pz with the factors as equal as possible (brick grid of the simulated ranks)."""
    best = (n, 1, 1)
    for px in range(1, n + 1):
        if n % px:
            continue
        for py in range(1, n // px + 1):
            if (n // px) % py:
                continue
            pz = n // px // py
            cand = tuple(sorted((px, py, pz), reverse=True))
            if max(cand) - min(cand) < max(best) - min(best):
                best = cand
    return best


class CpuWorld:
    """ONE global box of n^3 cubes, split into `ranks` bricks (the product's own partition rule: lower brick owns the
    shared planes, GhostMode::none).  Every simulated rank is one host thread: it assembles its owned cells into its own
    MatrixCSR (rows of shared vertices are ghost rows on the non-owners), sets the Dirichlet diagonal, and the restated
    MatrixCSR::scatter_rev (la/MatrixCSR.h:399-468) then adds the ghost rows to their owners - what `mpirun -np ranks`
    of the reference does, with the MPI exchange replaced by memory copies (which favours the CPU)."""

    def __init__(self, cfg, n, ranks):
        from oracle import oracle as O

        O.build(fast=True)
        M = _mesh_fixture()
        self.O, self.cfg, self.n, self.ranks = O, cfg, n, ranks
        pg = _grid3(ranks)
        nglob = (n, n, n)
        self.pgrid = pg
        self.bs = 3 if cfg == "q1" else 1
        ins = []
        for r in range(ranks):
            part = M.BoxPartition(r, pg, nglob)
            x = part.coordinates()
            if cfg == "p2":
                xd, dm, n_owned, ghosts, owners, dc = M.p2_partition(part)
            else:
                xd, dm, n_owned, ghosts, owners = part.tets() if cfg == "p1" else part.hexes()
                local = part.vertex_numbering()[0].reshape(-1)
                dc = np.empty((n_owned + len(ghosts), 3))
                dc[local] = x
            ins.append(dict(x=x, x_dofmap=xd, dofmap=dm, n_owned=n_owned, ghosts=ghosts, owners=owners, dc=dc))
        self.ins = ins
        maps = O.make_index_maps([i["n_owned"] for i in ins], [i["ghosts"] for i in ins], [i["owners"] for i in ins])
        rows, cols = [], []
        for i in ins:
            r_, c_ = O.sparsity_insert_cells(np.arange(len(i["dofmap"])), i["dofmap"], i["dofmap"])
            rows.append(r_)
            cols.append(c_)
        pats = O.sparsity_finalize(maps, maps, (self.bs, self.bs), rows, cols)
        del rows, cols
        self.mats = O.make_matrices(pats)
        self.maps = maps
        self.kid = {"p1": O.K_POISSON_P1_TET_A, "p2": O.K_POISSON_P2_TET_A, "q1": O.K_ELASTICITY_Q1_HEX_A}[cfg]
        self.consts = np.array([2.0]) if cfg != "q1" else np.array([1.0e9 / 2.6, 1.0e9 * 0.3 / (1.3 * 0.4)])
        self.markers, self.bc_rows, self.cells = [], [], []
        for i in ins:
            dc = i["dc"]
            mask = ((dc[:, 0] < 1e-12) | (dc[:, 1] > 1 - 1e-12)) if cfg == "q1" else ((dc[:, 0] < 1e-12) | (dc[:, 0] > 1 - 1e-12))
            bd = np.flatnonzero(mask).astype(np.int32)
            mk = np.zeros(len(dc) * self.bs, dtype=np.int8)
            O.bc_mark(mk, O.unroll_dofs(bd, self.bs))
            self.markers.append(mk)
            self.bc_rows.append(O.unroll_dofs(bd[bd < i["n_owned"]], self.bs))  # set_diagonal: owned rows only
            self.cells.append(np.arange(len(i["dofmap"]), dtype=np.int32))
        # vectorised form of the ghost-row exchange plan the oracle built (pack order: ghost-row order per neighbour)
        bs2 = self.bs * self.bs
        self.pack_idx, self.unpack_idx, self.ghost_begin = [], [], []
        for A in self.mats:
            m0 = A.index_maps[0]
            per = [[] for _ in m0.src]
            for i, g in enumerate(A.ghost_row_to_rank):
                r0, r1 = A.row_ptr[m0.size_local + i] * bs2, A.row_ptr[m0.size_local + i + 1] * bs2
                per[g].append(np.arange(r0, r1, dtype=np.int64))
            self.pack_idx.append([np.concatenate(b) if b else np.zeros(0, dtype=np.int64) for b in per])
            self.unpack_idx.append((A.unpack_pos[:, None] * bs2 + np.arange(bs2)[None, :]).reshape(-1))
            self.ghost_begin.append(int(A.row_ptr[m0.size_local]) * bs2)
        self.dofs_global = maps[0].size_global * self.bs
        self.n_cells = sum(len(i["dofmap"]) for i in ins)

    def scatter_rev(self):
        send = [[A.data[ix] for ix in self.pack_idx[r]] for r, A in enumerate(self.mats)]
        for r, A in enumerate(self.mats):
            m0 = A.index_maps[0]
            parts = [send[s][list(self.mats[s].index_maps[0].src).index(r)] for s in m0.dest]
            if parts:
                np.add.at(A.data, self.unpack_idx[r], np.concatenate(parts))
            A.data[self.ghost_begin[r]:] = 0.0

    def step(self, repeats=1):
        O = self.O

        def work(r):
            i, A = self.ins[r], self.mats[r]
            for _ in range(repeats):
                A.data[:] = 0.0
                O.assemble_matrix(self.kid, i["x_dofmap"], i["x"], self.cells[r], i["dofmap"], self.bs, i["dofmap"], self.bs,
                                  A.data, A.cols, A.row_ptr, bc0=self.markers[r], bc1=self.markers[r], constants=self.consts,
                                  fast=True)
                O.set_diagonal(A.data, A.cols, A.row_ptr, self.bs, self.bs, self.bc_rows[r], 1.0)

        t0 = time.perf_counter()
        for _ in range(1):
            ths = [threading.Thread(target=work, args=(r,)) for r in range(self.ranks)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        self.scatter_rev()
        return time.perf_counter() - t0

    def describe(self):
        pg = self.pgrid
        return (f"{self.ranks} host threads = {self.ranks} simulated MPI ranks ({pg[0]}x{pg[1]}x{pg[2]} bricks) of ONE "
                f"{self.n}^3 box ({self.n_cells} cells, {self.dofs_global} DOFs): oracle assemble_matrix(bcs) + set_diagonal "
                f"per rank, then the restated scatter_rev; liboracle_fast.so")


def cpu_assembly_rate(cfg, n, ranks, min_seconds):
    """(DOF/s, seconds per step, description) of the CPU arm: steps are repeated until min_seconds have passed."""
    w = CpuWorld(cfg, n, ranks)
    w.step()  # warm-up (page faults, thread start)
    times = []
    t_all = time.perf_counter()
    while time.perf_counter() - t_all < min_seconds or len(times) < 2:
        times.append(w.step())
    dt = float(np.median(times))
    return w.dofs_global / dt, dt, w.describe() + f"; median of {len(times)} steps", w


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


CPU_SAMPLE_N = {"p1": 64, "p2": 32, "q1": 48}      # cpu_baseline leg of the GPU arm (about 10 s of CPU work + setup)
CPU_REFERENCE_N = {"p1": 128, "p2": 48, "q1": 64}   # --impl reference: the largest box whose Python-side setup stays short


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n = args.n_cpu or CPU_REFERENCE_N[args.config]
    t0 = time.perf_counter()
    w = CpuWorld(args.config, n, cores)
    t_setup = time.perf_counter() - t0
    for _ in range(args.warmup):
        w.step()
    times = [w.step() for _ in range(args.steps)]
    dt = float(np.mean(times))
    value = w.dofs_global / dt
    sample = w.describe() + f"; setup {t_setup:.0f} s (untimed)"
    # the line carries the GPU arm's config (task statement, section 4); what was actually assembled - a bounded sample of
    # that workload - is named in cpu_baseline.sample and under "sample"
    cfgd = workload_config(args.config, args.n or CONFIGS[args.config][2], args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfgd,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample": {"cells_per_edge": n, "dofs": w.dofs_global, "cells": w.n_cells, "simulated_ranks": cores},
    }
    emit(line)


def workload_config(cfg, n, world):
    names = {"p1": "Poisson P1 tets", "p2": "Poisson P2 tets", "q1": "Linear elasticity Q1 hexes (bs=3)"}
    pg = _mesh_fixture().pgrid_for(world) if world > 1 else (1, 1, 1)
    return {"workload": f"{names[cfg]} on a {n * pg[0]}x{n * pg[1]}x{n * pg[2]} box: assemble_matrix(bcs) + "
                        "set_diagonal + scatter_reverse into MatrixCSR",
            "cells_per_gpu": n, "bricks": list(pg), "l2": "inputs larger than L2 (no flush needed)"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
_RESULT_FD = None


def quiet_stdout():
    """Libraries (NCCL prints its version) must not add lines to stdout: everything but the result line goes to
    stderr; emit() writes the one JSON line to the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


class Ctx:
    """What every leg needs: torch / torch.distributed, the product modules, ranks, device, communicator."""


def kernel_ids(K, elem):
    if elem == "P1":
        return K.K_POISSON_P1_TET_A, K.K_LOAD_P1_TET_L
    if elem == "P2":
        return K.K_POISSON_P2_TET_A, K.K_LOAD_P2_TET_L
    return K.K_ELASTICITY_Q1_HEX_A, K.K_LOAD_Q1_HEX_L


def oracle_kernel(O, cfg):
    if cfg == "p1":
        return O.K_POISSON_P1_TET_A, np.array([2.0])
    if cfg == "p2":
        return O.K_POISSON_P2_TET_A, np.array([2.0])
    return O.K_ELASTICITY_Q1_HEX_A, np.array([1.0e9 / 2.6, 1.0e9 * 0.3 / (1.3 * 0.4)])


PARITY_TOL = 1e-12


def nullspace_residual(ctx, pb, A, assemble):
    """max |A m| / (max |A| max |m|) over the constant modes (Poisson: 1, cpp/test/matrix.cpp:96-109; elasticity: the
    three translations) of the matrix assembled WITHOUT bcs: a size-independent check of assembly + scatter_rev + mult
    (with its forward scatter) that costs a few launches at any size and on any number of GPUs."""
    torch, la = ctx.torch, ctx.la
    bs = pb["bs"]
    A.set_value(0.0)
    assemble(A, [])
    amax = float(A._values().abs().max())
    if ctx.world > 1:
        amax = ctx.comm.allreduce_max(amax)
    x = la.Vector(A.index_map(1), bs)
    y = la.Vector(A.index_map(0), bs)
    worst = 0.0
    for k in range(bs):
        x.array.zero_()
        x.array.view(-1, bs)[:, k] = 1.0
        y.set(0.0)
        A.mult(x, y)
        r = float(y.array[: bs * A.num_owned_rows()].abs().max())
        worst = max(worst, ctx.comm.allreduce_max(r) if ctx.world > 1 else r)
    return worst / amax


def parity_single(ctx, cfg, pb, A, assemble):
    """N = 1: the assembled matrix (bcs applied, diagonal set) against the CPU oracle on ~500 sampled rows incl. rows on
    the Dirichlet boundary (tests/sampled_rows.py, the check of tests/test_gpu_fullsize.py), + the null-space residual.
    The oracle is the checker here; nothing in the timed legs touches it."""
    from oracle import oracle as O
    from tests import sampled_rows

    torch, fem = ctx.torch, ctx.fem
    O.build()
    V, bc, bs = pb["V"], pb["bc"], pb["bs"]
    A.set_value(0.0)
    assemble(A, [bc])
    fem.set_diagonal(A, V, [bc], 1.0)
    mk = fem._bc_markers(V, [bc]).cpu().numpy()
    rng = np.random.default_rng(7)
    ndofs = pb["ndofs_local"]
    bdofs = bc._dofs0[:: max(1, len(bc._dofs0) // 100)] // bs
    rows = np.unique(np.concatenate([rng.integers(0, ndofs, 400), bdofs, [0, ndofs - 1]])).astype(np.int64)
    kid, consts = oracle_kernel(O, cfg)
    t0 = time.perf_counter()
    ud, pat, ref, ncells_sub = sampled_rows.oracle_rows(torch, O, pb, kid, consts, mk, rows)
    err = sampled_rows.compare_rows(A, rows, ud, pat, ref, bs)
    res = nullspace_residual(ctx, pb, A, assemble)
    return {"kind": "sampled rows vs CPU oracle (values row-scaled, columns exact) + null space of the bc-free matrix",
            "rows_sampled": int(len(rows)), "oracle_cells": ncells_sub, "max_rel_err_vs_oracle": err,
            "nullspace_residual": res, "tol": PARITY_TOL, "ok": bool(err <= PARITY_TOL and res <= 64 * PARITY_TOL),
            "seconds": time.perf_counter() - t0}


def parity_multi(ctx, cfg, pb, A, assemble):
    """N > 1: (i) null-space residual of the full-size distributed matrix; (ii) the reference's own serial-vs-parallel
    check (cpp/test/matrix.cpp:59-64): |A|_F^2 and |b|_2 (load vector, lifted with a non-zero g, ghost contributions
    added) of a small box assembled on the N ranks against the same box assembled on one rank (every rank repeats the
    serial assembly on its own GPU with a serial communicator)."""
    torch, fem, la, common = ctx.torch, ctx.fem, ctx.la, ctx.common
    t0 = time.perf_counter()
    res = nullspace_residual(ctx, pb, A, assemble)
    n_small = {"p1": 24, "p2": 12, "q1": 16}[cfg]
    from dolfinx_b200 import mesh as M

    pg = M.pgrid_for(ctx.world)

    def norms(comm, n):
        q = build_problem(cfg, n, comm, ctx.device)
        sp = fem.create_sparsity_pattern(q["a"])
        sp.finalize()
        B = la.MatrixCSR(sp)
        if comm.size > 1:
            fem.assemble_matrix_overlapped(B, q["a"], bcs=[q["bc"]])
        else:
            fem.assemble_matrix(B, q["a"], bcs=[q["bc"]])
            B.scatter_reverse()
        fem.set_diagonal(B, q["V"], [q["bc"]], 1.0)
        b = la.Vector(q["V"].dofmap.index_map, q["bs"])
        fem.assemble_vector(b, q["L"])
        fem.apply_lifting(b, [q["a"]], [[q["bc"]]])
        b.scatter_reverse(la.InsertMode.add)
        fem.set_bc(b, [q["bc"]])
        return B.squared_norm(), la.norm(b)

    a_par, b_par = norms(ctx.comm, n_small)
    a_ser, b_ser = norms(common.COMM_SELF, tuple(n_small * pg[d] for d in range(3)))
    da, db = abs(a_par - a_ser) / a_ser, abs(b_par - b_ser) / b_ser
    worst = ctx.comm.allreduce_max(max(da, db))
    return {"kind": "null space of the full-size matrix + |A|_F^2, |b|_2 of a small box on N ranks vs one rank",
            "nullspace_residual": res, "small_box": [n_small * pg[d] for d in range(3)],
            "normA2_parallel": a_par, "normA2_serial": a_ser, "normb_parallel": b_par, "normb_serial": b_ser,
            "max_rel_diff": worst, "tol": PARITY_TOL, "ok": bool(worst <= PARITY_TOL and res <= 64 * PARITY_TOL),
            "seconds": time.perf_counter() - t0}


def run_leg(ctx, args, cfg, n, headline, lean=False, steps=None):
    """One BASELINE config on the N ranks: timed step loop, dominant kernel alone, load vector, lifting, SpMV, parity;
    the headline leg adds the end-to-end (host buffer) loop.  Returns the result dict on rank 0 (None elsewhere)."""
    torch, dist, K, fem, la = ctx.torch, ctx.dist, ctx.K, ctx.fem, ctx.la
    world, rank, device, comm = ctx.world, ctx.rank, ctx.device, ctx.comm
    steps = steps or args.steps
    cell, elem, n_def = CONFIGS[cfg]
    torch.cuda.reset_peak_memory_stats()

    t0 = time.perf_counter()
    pb = build_problem(cfg, n, comm, device)
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t0
    a, L, V, bc = pb["a"], pb["L"], pb["V"], pb["bc"]

    t0 = time.perf_counter()
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)
    torch.cuda.synchronize()
    t_pattern = time.perf_counter() - t0
    strat = {"auto": None, "atomic": K.ASM_ATOMIC, "chunked": K.ASM_CHUNKED, "rowgather": K.ASM_ROWGATHER}[args.strategy]
    t0 = time.perf_counter()
    integ0 = a.integral(fem.IntegralType.cell, 0)
    subset_plans = False
    if lean:
        # near the HBM limit: build only what the distributed step needs.  Lean P1 plans run ONE plan of all cells in two
        # parts; everything else runs plans over the boundary / interior cell subsets, and the kernel-only leg then
        # times the interior launch instead of building a third plan
        fem.assemble_matrix_overlapped(A, a, bcs=[bc], strategy=strat)
        subset_plans = any(isinstance(k, tuple) and k[0] == "plan" and k[3] == "int" for k in a._plans)
    if subset_plans:
        cells_int = fem._boundary_interior_cells(a, integ0)[1]
        kplan = fem._asm_plan(a, integ0, fem.IntegralType.cell, A, subset=("int", cells_int))
        strat_used = fem._matrix_strategy(a, integ0, kplan, strat, shared=True)
        kernel_cells = int(cells_int.numel())
    else:
        fem.assemble_matrix(A, a, bcs=[bc], strategy=strat)  # builds the assembly plan (and its chunk lists)
        kplan = fem._asm_plan(a, integ0, fem.IntegralType.cell, A)
        # the strategy the product path actually runs on the whole cell list (what the kernel-only and e2e legs time)
        strat_used = fem._matrix_strategy(a, integ0, kplan, strat)
        kernel_cells = pb["n_cells"]
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    strat_name = {K.ASM_ATOMIC: "atomic", K.ASM_CHUNKED: "chunked", K.ASM_ROWGATHER: "rowgather"}[strat_used]

    plan_info = None
    if strat_used == K.ASM_CHUNKED and world == 1:
        nch, ndest, nsrc, pbytes = fem.chunk_stats(a, A)
        plan_info = {"chunks": nch, "destinations": ndest, "list_entries": nsrc, "plan_bytes": pbytes,
                     "plan_bytes_per_cell": pbytes / pb["n_cells"]}
    nnz = A._nnz
    n_rows = A.num_all_rows()
    n_cols = A.index_map(1).size_local + A.index_map(1).num_ghosts
    b_asm = alg_bytes_asm(pb, nnz, n_rows) * kernel_cells // pb["n_cells"]  # (lean: the interior launch's share)
    b_spmv = alg_bytes_spmv(pb, nnz, A.num_owned_rows(), n_cols)
    hbm_peak, peak_src = peaks()

    def assemble(M_, bcs_):
        if world > 1:
            # boundary cells -> ghost-row exchange on the comm stream -> interior cells -> add received rows
            fem.assemble_matrix_overlapped(M_, a, bcs=bcs_, strategy=strat)
        else:
            fem.assemble_matrix(M_, a, bcs=bcs_, strategy=strat)
            M_.scatter_reverse()

    def step():
        # one time-step's worth of the hot path: A <- 0, assemble, ghost-row exchange, diagonal.  The zero-fill is
        # fused into kernels that write every value once (row-gather) and is a real memset otherwise.
        A.set_value(0.0)
        assemble(A, [bc])
        fem.set_diagonal(A, V, [bc], 1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    bc0 = fem._bc_markers(V, [bc])
    carr, ncst = K.constants_array(fem.pack_constants(a))
    cf = K.make_coeffs()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for i in range(steps):
        step()
    e_stop.record()
    barrier()
    ms_total = e_start.elapsed_time(e_stop)
    # dominant-kernel launches alone, same stream, CUDA events around every launch
    # (same values mode as the step: the aggregated kernels run on a zeroed matrix in overwrite mode)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kmode = K.VALUES_ADD if strat_used == K.ASM_ATOMIC else K.VALUES_OVERWRITE
    for i in range(steps):
        A.set_value(0.0)
        vals = A._values()  # zero-fill written here, outside the event pair
        ev[i][0].record()
        K.check(K.lib.bfx_assemble_matrix_cells(kplan, integ0.kernel, a.mesh.x.data_ptr(), bc0.data_ptr(), bc0.data_ptr(),
                                                C.byref(cf), carr, ncst, vals.data_ptr(), strat_used, kmode,
                                                K.current_stream()))
        ev[i][1].record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_kernel = float(np.mean([s.elapsed_time(e) for s, e in ev]))

    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / steps
    value = pb["ndofs_global"] * pb["bs"] / (ms_step * 1e-3)

    # ---- vector assembly, lifting (non-zero g), SpMV (second timed loops) ------------------------------
    b = la.Vector(V.dofmap.index_map, pb["bs"])
    fem.assemble_vector(b, L)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(5):
        fem.assemble_vector(b, L)
        b.scatter_reverse(la.InsertMode.add)
    e1.record()
    barrier()
    ms_vec = e0.elapsed_time(e1) / 5
    # apply_lifting (fem/assembler.h:336-493) for the same bc (g = 1, 2, .. per component): b <- b - A g
    fem.apply_lifting(b, [a], [[bc]])
    barrier()
    e0.record()
    for _ in range(5):
        fem.apply_lifting(b, [a], [[bc]])
    e1.record()
    barrier()
    ms_lift = e0.elapsed_time(e1) / 5
    step()
    if args.spmv_variant != -1:
        K.check(K.lib.bfx_csr_set_spmv_variant(A._csr, args.spmv_variant))
    x = la.Vector(A.index_map(1), pb["bs"])
    y = la.Vector(A.index_map(0), pb["bs"])
    g = torch.Generator(device=device)
    g.manual_seed(12345)
    x.array.copy_(torch.rand(x.array.numel(), generator=g, device=device, dtype=torch.float64))
    for _ in range(10):
        A.mult(x, y)
    barrier()
    e0.record()
    for _ in range(args.spmv_reps):
        A.mult(x, y)
    e1.record()
    barrier()
    ms_spmv = e0.elapsed_time(e1) / args.spmv_reps
    tt = torch.tensor([ms_spmv, ms_vec, ms_lift, ms_kernel], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_spmv, ms_vec, ms_lift, ms_kernel = (float(v) for v in tt)
    del x, y, b

    # ---- phases of one distributed step (CUDA events on the launching stream, mean of 5 steps, this rank) -------
    timeline = None
    if world > 1 and args.timeline:
        acc = {}
        for _ in range(5):
            barrier()
            tl = []
            A.set_value(0.0)
            fem.assemble_matrix_overlapped(A, a, bcs=[bc], strategy=strat, timeline=tl)
            ev_d = torch.cuda.Event(enable_timing=True)
            fem.set_diagonal(A, V, [bc], 1.0)
            ev_d.record()
            tl.append(("set_diagonal", ev_d))
            torch.cuda.synchronize()
            for (n0, e0_), (n1, e1_) in zip(tl[:-1], tl[1:]):
                acc.setdefault(n1, []).append(e0_.elapsed_time(e1_))
        timeline = {k: float(np.mean(v)) for k, v in acc.items()}
        tl_all = [None] * world
        dist.all_gather_object(tl_all, timeline)
        timeline = {"per_rank_ms": tl_all, "note": "gaps between events on the launching stream; the exchange itself runs "
                    "on the communication stream and shows up as waiting time in 'wait + unpack'"}

    # ---- end-to-end through the host-buffer C-ABI entry (headline leg) ---------------------------------
    e2e = None
    if headline and not args.no_e2e and not lean:
        e2e = run_e2e(ctx, args, pb, A, kplan, integ0.kernel, bc0, carr, ncst, strat_used, barrier)

    # ---- parity (checker; untimed) ----------------------------------------------------------------------
    parity = None
    if not args.no_parity and not lean:
        parity = parity_single(ctx, cfg, pb, A, assemble) if world == 1 else parity_multi(ctx, cfg, pb, A, assemble)

    out = None
    if rank == 0:
        ach = b_asm / (ms_kernel * 1e-3) / 1e9
        out = {
            "value": value, "ms_per_step": ms_step, "steps": steps,
            "workload": workload_config(cfg, n, world)["workload"],
            "roofline": {"bound": "hbm", "kernel": f"assemble_cells_matrix[{cfg},{strat_name}]",
                         "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": measured_traffic(cfg, strat_name, n),
                         "peak_source": peak_src, "alg_bytes_per_launch": b_asm, "kernel_ms": ms_kernel,
                         "kernel_cells": kernel_cells},
            "e2e": e2e,
            "parity": parity,
            "clocks": clocks,
            "spmv": {"ms": ms_spmv, "gbs": b_spmv / (ms_spmv * 1e-3) / 1e9, "frac": b_spmv / (ms_spmv * 1e-3) / 1e9 / hbm_peak,
                     "gflops": 2 * pb["bs"] ** 2 * nnz / (ms_spmv * 1e-3) / 1e9, "alg_bytes": b_spmv, "reps": args.spmv_reps},
            "vector_assembly_ms": ms_vec,
            "apply_lifting_ms": ms_lift,
            "sizes": {"dofs_global": pb["ndofs_global"] * pb["bs"], "cells_per_gpu": pb["n_cells"], "nnz_per_gpu": nnz},
            "setup_s": {"mesh": t_mesh, "sparsity+matrix": t_pattern, "assembly_plan+first_call": t_plan},
            "chunk_plan": plan_info,
            "timeline": timeline,
            "hbm": {"peak_allocated_gb": torch.cuda.max_memory_allocated() / 1e9,
                    "free_gb_at_end": torch.cuda.mem_get_info()[0] / 1e9, "total_gb": torch.cuda.mem_get_info()[1] / 1e9},
        }
    return out


def run_e2e(ctx, args, pb, A, plan, kid, bc0, carr, ncst, strat_used, barrier):
    torch, dist, K = ctx.torch, ctx.dist, ctx.K
    a = pb["a"]
    nval = A._values().numel()
    x_host = torch.empty((pb["n_x"], 3), dtype=torch.float64, pin_memory=True)
    x_host.copy_(a.mesh.x)
    mk_host = torch.empty(bc0.numel(), dtype=torch.int8, pin_memory=True)
    mk_host.copy_(bc0)
    out_host = torch.empty(nval, dtype=torch.float64, pin_memory=True)
    reps = max(2, min(args.steps, 6))
    out_host2 = torch.empty(nval, dtype=torch.float64, pin_memory=True)

    def e2e_step():
        K.check(K.lib.bfx_assemble_matrix_cells_host(
            plan, kid, x_host.data_ptr(), pb["n_x"], mk_host.data_ptr(),
            mk_host.data_ptr(), mk_host.numel(), None, 0, 1, carr, ncst, out_host.data_ptr(), strat_used, K.current_stream()))

    def e2e_begin(out):
        K.check(K.lib.bfx_assemble_matrix_cells_host_begin(
            plan, kid, x_host.data_ptr(), pb["n_x"], mk_host.data_ptr(),
            mk_host.data_ptr(), mk_host.numel(), None, 0, 1, carr, ncst, out.data_ptr(), strat_used))

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        td = torch.tensor([dt], dtype=torch.float64, device=ctx.device)
        if ctx.world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        return float(td.item())

    def run_sync():
        for _ in range(reps):
            e2e_step()

    def run_pipelined():
        # the time loop a caller writes with the split entry: step k+1 is enqueued before step k is waited
        # for; every step still uploads its inputs and downloads all its values into pinned host memory
        outs = (out_host, out_host2)
        e2e_begin(outs[0])
        for k in range(1, reps):
            e2e_begin(outs[k & 1])
            K.check(K.lib.bfx_assemble_matrix_cells_host_end(plan))
        K.check(K.lib.bfx_assemble_matrix_cells_host_end(plan))

    e2e_step()
    run_pipelined()
    dt_sync = timed(run_sync)
    dt = timed(run_pipelined)
    err2 = float((out_host - out_host2).abs().max() / out_host.abs().max())  # the two slots hold the same matrix
    dofs = pb["ndofs_global"] * pb["bs"]
    return {"value": dofs / dt, "unit": UNIT,
            "h2d_bytes_per_step": int(x_host.numel() * 8 + mk_host.numel()), "d2h_bytes_per_step": int(nval * 8),
            "ms_per_step": dt * 1e3,
            "call": "bfx_assemble_matrix_cells_host_begin/_end, two steps in flight (pinned host buffers)",
            "one_step_in_flight": {"value": dofs / dt_sync, "ms_per_step": dt_sync * 1e3,
                                   "call": "bfx_assemble_matrix_cells_host"},
            "slots_max_rel_diff": err2}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="bfx", choices=["bfx", "reference"])
    ap.add_argument("--config", default="p1", choices=list(CONFIGS))
    # (--cells-per-edge: the spelling to use under torchrun, whose own parser rejects "--n" as ambiguous)
    ap.add_argument("--n", "--cells-per-edge", dest="n", type=int, default=0,
                    help="cells per box edge (per GPU); 0 = BASELINE size")
    ap.add_argument("--strategy", default="auto", choices=["auto", "atomic", "chunked", "rowgather"],
                    help="scatter-add strategy of the matrix kernel; auto = the aggregated kernel of the element")
    ap.add_argument("--spmv-reps", type=int, default=100)
    ap.add_argument("--spmv-variant", type=int, default=-1,
                    help="bs=1 SpMV kernel (0 stream, 1 rows, 2 TMA rows); -1 = by row length, -2 = timed selection")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--n-cpu", type=int, default=0, help="cells per edge of the CPU arm's box (0 = default per config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="N > 1: per-phase CUDA-event times of the distributed step")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity legs (oracle on sampled rows / norms)")
    ap.add_argument("--legs", default="auto",
                    help="other BASELINE configs measured in the same run and reported under 'configs': comma list of "
                         "p2,q1,c5 | none | auto (default run of the headline config: p2,q1; and c5 = P1 500^3 per GPU "
                         "on 8 GPUs)")
    ap.add_argument("--c5-n", type=int, default=500, help="cells per edge and GPU of the c5 leg")
    ap.add_argument("--lean", action="store_true",
                    help="N > 1 at shard sizes near the HBM limit (C5: --n 500): never build the plan of the whole "
                         "cell list beside the boundary / interior plans of the overlapped assembly; implies --no-e2e")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "bfx" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import gc

    import torch
    import torch.distributed as dist

    from dolfinx_b200 import _lib as K
    from dolfinx_b200 import common, fem, la

    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.K, ctx.common, ctx.fem, ctx.la = torch, dist, K, common, fem, la
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(ctx.local_rank)
    ctx.device = torch.device("cuda", ctx.local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=ctx.device)
    ctx.comm = common.Comm()
    n = args.n or CONFIGS[args.config][2]
    lean = args.lean and world > 1

    def release():
        gc.collect()
        torch.cuda.empty_cache()

    head = run_leg(ctx, args, args.config, n, headline=True, lean=lean)
    release()

    legs = {}
    if args.legs == "auto":
        names = []
        if args.config == "p1" and not args.n and not lean:
            names = ["p2", "q1"] + (["c5"] if world == 8 else [])
    else:
        names = [s for s in args.legs.split(",") if s and s != "none"]
    for name in names:
        # (a leg that fails - e.g. a shard that does not fit - must not take the headline line with it; the failure is
        # agreed on by all ranks so that nobody waits in a collective of a leg the others left)
        leg, err = None, None
        try:
            if name == "c5":
                if world == 1:
                    continue
                leg = run_leg(ctx, args, "p1", args.c5_n, headline=False, lean=True, steps=min(args.steps, 10))
            else:
                leg = run_leg(ctx, args, name, CONFIGS[name][2], headline=False, steps=min(args.steps, 10))
        except Exception as e:  # noqa: BLE001
            import traceback

            where = " <- ".join(f"{f.name}:{f.lineno}" for f in traceback.extract_tb(e.__traceback__)[-4:])
            err = (f"{type(e).__name__}: {e}"[:300]) + " @ " + where
        release()
        failed = torch.tensor([1.0 if err else 0.0], device=ctx.device)
        if world > 1:
            dist.all_reduce(failed, op=dist.ReduceOp.MAX)
        if rank == 0:
            legs[name] = leg if float(failed.item()) == 0.0 else {"error": err or "failed on another rank"}
        if float(failed.item()) != 0.0 and world > 1:
            break  # ranks may be out of step inside the failed leg's collectives: no further legs

    cpu = None
    if rank == 0 and not args.no_cpu:
        cores = host_cores()
        ncpu = args.n_cpu or CPU_SAMPLE_N[args.config]
        rate, dtc, desc, _w = cpu_assembly_rate(args.config, ncpu, cores, 8.0)
        del _w
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc + f" ({dtc * 1e3:.1f} ms per step)"}

    if rank == 0:
        # kernels of libbfx launched inside the timed step loop: marker bit-packing + the assembly kernel (two launches on N > 1 GPUs:
        # boundary cells, interior cells) + set_diagonal, and on N > 1 the pack / unpack kernels of scatter_rev
        per_step = 3 if world == 1 else 6
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.config, n, world),
            "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": head["e2e"], "parity": head["parity"],
            "gpu_launches": per_step * args.steps, "clocks": head["clocks"],
            "spmv": head["spmv"], "vector_assembly_ms": head["vector_assembly_ms"],
            "apply_lifting_ms": head["apply_lifting_ms"], "sizes": head["sizes"], "setup_s": head["setup_s"],
            "chunk_plan": head["chunk_plan"], "timeline": head.get("timeline"), "hbm": head["hbm"],
            "configs": legs,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
