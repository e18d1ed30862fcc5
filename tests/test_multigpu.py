"""Row (e) on real devices: two ranks (NCCL), direction shards + cell-sharded post-J stages, against the single-GPU solve.
Needs 2 GPUs (skipped otherwise); the host-side logic of the same path is covered on CPU by the gloo test."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(V):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_grid
    from voronoirt_b200 import synth
    pos, nbr, b = load_grid("grid_strat3000")
    n = pos.shape[1]
    a = synth.atmosphere(pos[0], pos[1], pos[2])
    cell = V.read_cell(nbr, n, pos, b[2], b[3], b[4], b[5])
    sites = V.VoronoiSites(*cell, a["temperature"], a["electron_density"], a["hydrogen_density"], a["velocity_z"],
                           a["velocity_x"], a["velocity_y"], b[0], b[1], b[2], b[3], b[4], b[5], n)
    line, lte, α_cont, ελ, Cr = synth.line_inputs(a["temperature"], a["electron_density"], a["hydrogen_density"], 10, 4)
    return sites, line, lte, α_cont, ελ, Cr


def _ferry_id(dist, torch, api, rank, src=0):
    """what a host does for vrt_solver_comm_init: one process makes the NCCL unique id, the 128 bytes travel to the others"""
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == src:
        buf = torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy().tobytes()


def _attach_peers(dist, torch, s, world):
    """J reduced through peer memory: every rank's 64-byte IPC handle, gathered in rank order"""
    mine = torch.frombuffer(bytearray(s.peer_handle()), dtype=torch.uint8).cuda()
    allh = torch.zeros(64 * world, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(allh, mine)
    s.peer_attach(allh.cpu().numpy().tobytes())


def _worker(rank, world, port, q, mode):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib, api
    import bench
    _lib.check(_lib.lib().vrt_set_device(rank))
    qp = V.quadrature_path("ul7n12")
    if mode == "regular":
        # direction shards on the regular grid (Λ_regular, lambda_iteration.jl:116-205): J is all-reduced inside the library
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        from regular_box import regular_line_box
        P = regular_line_box(O)
        f = P["fields"]
        atm = V.Atmosphere(P["z"], P["x"], P["y"], f["temperature"], f["electron_density"], f["hydrogen_density"], f["velocity_z"],
                           f["velocity_x"], f["velocity_y"])
        dlo, dhi = bench.shard_range(12, world, rank)
        s = V.Solver(atm, qp, line=P["line"], α_cont=P["α_cont"], ελ=P["ελ"], C_rates=P["C"], LTE_pops=P["lte"], dir_range=(dlo, dhi))
        s.comm_init(dir_id=_ferry_id(dist, torch, api, rank), dir_rank=rank, dir_size=world)
        res = s.iterate(-1.0, 2)
        S, J, pops = s.get_state()
        q.put((rank, dict(S=S, J=J, pops=pops, diffs=[h["diff"] for h in res["history"]])))
        dist.barrier()
        s.close()
        dist.destroy_process_group()
        return
    sites, line, lte, α_cont, ελ, Cr = _problem(V)
    out = {}
    if mode == "lambda":
        # wavelength shards (BASELINE configs[2]): every rank all directions, a slice of the wavelengths; the rates are summed
        # over the wavelength group inside the library
        lo, hi = bench.shard_range(len(line.λ), world, rank)
        s = V.Solver(sites, qp, line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte, lam_range=(lo, hi))
        s.comm_init(lam_id=_ferry_id(dist, torch, api, rank), lam_rank=rank, lam_size=world)
        res = s.iterate(-1.0, 3)
        S, J, pops = s.get_state()
        out = dict(S=S, J=J, pops=pops, lo=lo, hi=hi, diffs=[h["diff"] for h in res["history"]], checksum=s.checksum())
        q.put((rank, out))
        dist.barrier()
        s.close()
        dist.destroy_process_group()
        return
    dlo, dhi = bench.shard_range(12, world, rank)
    s = V.Solver(sites, qp, line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte, dir_range=(dlo, dhi), cell_shard=(rank, world))
    if mode in ("nccl", "peers"):
        assert _lib.lib().vrt_nccl_available() == 1
        s.comm_init(dir_id=_ferry_id(dist, torch, api, rank), dir_rank=rank, dir_size=world)
        if mode == "peers":
            _attach_peers(dist, torch, s, world)
    else:
        class _Dev:
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}

        def hook(ptr, count, op):
            if op == 0:
                return 0                     # sum over wavelength shards: there is only one here, the rates are already complete
            t = torch.as_tensor(_Dev(ptr, count), device=torch.device("cuda", rank))
            assert t.data_ptr() == ptr       # a view of the library's buffer, not a copy
            if op in (3, 4):
                sl = t[rank * (count // world):(rank + 1) * (count // world)]
                if op == 3:
                    dist.reduce_scatter_tensor(sl, t)
                else:
                    dist.all_gather_into_tensor(t, sl)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            return 0
        s.set_allreduce(hook)
    res = s.iterate(-1.0, 2)
    # cell-sliced checkpoint round trip: every rank takes its own slice out, puts it back in (the other slices arrive through
    # the all-gather), then one more iteration
    c0, c1 = s.cell_slice()
    Ss, Js, ps = s.get_state_slice()
    s.set_state_slice(Ss, ps)
    res2 = s.iterate(-1.0, 1)
    chk = s.checksum()
    S, J, pops = s.get_state()
    out = dict(S=S, J=J, pops=pops, diffs=[h["diff"] for h in res["history"]] + [h["diff"] for h in res2["history"]], slice=(c0, c1),
               checksum=chk, S_slice=Ss)
    q.put((rank, out))
    s.peer_detach()          # (no-op unless peers were attached) importers first, then a barrier, then the solvers go
    dist.barrier()
    s.close()
    dist.destroy_process_group()


def _run(mode, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + hash(mode)) % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    return got


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("mode", ["hook", "nccl", "peers"])
def test_two_rank_direction_shards_match_single_gpu(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import voronoirt_b200 as V
    sites, line, lte, α_cont, ελ, Cr = _problem(V)
    ref = V.Solver(sites, V.quadrature_path("ul7n12"), line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte)
    rres = ref.iterate(-1.0, 3)
    S1, J1, p1 = ref.get_state()
    chk1 = ref.checksum()
    ref.close()
    got = _run(mode)
    r0, r1 = got[0], got[1]
    for r in (r0, r1):                                   # both ranks hold the full, identical state
        assert rel(r["S"], S1) < 1e-12 and rel(r["J"], J1) < 1e-12
        assert np.all(np.abs(r["pops"] - p1) <= 1e-9 * np.abs(p1) + 1e-13 * sites.hydrogen_populations[:, None])
        # two iterations, the cell-sliced checkpoint round trip, one more iteration: the criterion of a fresh
        # vrt_lambda_iterate call starts from S_old = 0 (lambda_iteration.jl:241), so its first value is 1
        assert np.allclose(r["diffs"][:2], [h["diff"] for h in rres["history"]][:2], rtol=1e-9) and r["diffs"][2] == 1.0
        assert abs(r["checksum"]["sum_S"] / chk1["sum_S"] - 1) < 1e-12 and abs(r["checksum"]["sum_populations"] / chk1["sum_populations"] - 1) < 1e-12
    assert np.array_equal(r0["S"], r1["S"])
    # the cell slices tile the cells and are rows of S in internal (perm_up) order
    n = sites.n
    assert r0["slice"][0] == 0 and r0["slice"][1] == r1["slice"][0] and r1["slice"][1] == n
    assert abs((r0["checksum"]["sum_J_own_cells"] + r1["checksum"]["sum_J_own_cells"]) / chk1["sum_J_own_cells"] - 1) < 1e-12


def test_two_rank_wavelength_shards_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import voronoirt_b200 as V
    sites, line, lte, α_cont, ελ, Cr = _problem(V)
    ref = V.Solver(sites, V.quadrature_path("ul7n12"), line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte)
    rres = ref.iterate(-1.0, 3)
    S1, J1, p1 = ref.get_state()
    ref.close()
    got = _run("lambda")
    for r in got.values():
        lo, hi = r["lo"], r["hi"]
        assert rel(r["S"], S1[lo:hi]) < 1e-12 and rel(r["J"], J1[lo:hi]) < 1e-12
        assert np.all(np.abs(r["pops"] - p1) <= 1e-9 * np.abs(p1) + 1e-13 * sites.hydrogen_populations[:, None])
        assert np.allclose(r["diffs"], [h["diff"] for h in rres["history"]], rtol=1e-9)


def test_two_rank_regular_grid_direction_shards():
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import voronoirt_b200 as V
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from regular_box import regular_line_box
    P = regular_line_box(O)
    f = P["fields"]
    atm = V.Atmosphere(P["z"], P["x"], P["y"], f["temperature"], f["electron_density"], f["hydrogen_density"], f["velocity_z"],
                       f["velocity_x"], f["velocity_y"])
    ref = V.Solver(atm, V.quadrature_path("ul7n12"), line=P["line"], α_cont=P["α_cont"], ελ=P["ελ"], C_rates=P["C"], LTE_pops=P["lte"])
    rres = ref.iterate(-1.0, 2)
    S1, J1, p1 = ref.get_state()
    ref.close()
    got = _run("regular")
    for r in got.values():
        # the shards add the directions in another order than the single-GPU layout grouping: rounding only
        assert rel(r["S"], S1) < 1e-12 and rel(r["J"], J1) < 1e-12
        assert np.allclose(r["diffs"], [h["diff"] for h in rres["history"]], rtol=1e-9)
        assert np.all(np.abs(r["pops"] - p1) <= 1e-9 * np.abs(p1) + 1e-13 * P["flat"]["hydrogen_density"][:, None])
