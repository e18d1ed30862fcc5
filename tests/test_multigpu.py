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


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import voronoirt_b200 as V
    from voronoirt_b200 import _lib
    import bench
    _lib.check(_lib.lib().vrt_set_device(rank))
    sites, line, lte, α_cont, ελ, Cr = _problem(V)
    qp = V.quadrature_path("ul7n12")
    dlo, dhi = bench.shard_range(12, world, rank)
    s = V.Solver(sites, qp, line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte, dir_range=(dlo, dhi), cell_shard=(rank, world))

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
    res = s.iterate(-1.0, 3)
    S, J, pops = s.get_state()
    if rank == 0:
        q.put((S, J, pops, [h["diff"] for h in res["history"]]))
    dist.barrier()
    s.close()
    dist.destroy_process_group()


def test_two_rank_direction_shards_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import voronoirt_b200 as V
    sites, line, lte, α_cont, ελ, Cr = _problem(V)
    ref = V.Solver(sites, V.quadrature_path("ul7n12"), line=line, α_cont=α_cont, ελ=ελ, C_rates=Cr, LTE_pops=lte)
    rres = ref.iterate(-1.0, 3)
    S1, J1, p1 = ref.get_state()
    ref.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    S2, J2, p2, diffs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)

    def rel(a, b):
        return np.abs(a - b).max() / np.abs(b).max()
    assert rel(S2, S1) < 1e-12 and rel(J2, J1) < 1e-12
    assert np.all(np.abs(p2 - p1) <= 1e-9 * np.abs(p1) + 1e-13 * sites.hydrogen_populations[:, None])
    assert np.allclose(diffs, [h["diff"] for h in rres["history"]], rtol=1e-9)
