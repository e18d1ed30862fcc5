"""Host-side logic of the multi-GPU decomposition (bench.py): direction / wavelength shard bookkeeping and the
longest-processing-time-first balance, including a world_size-2 run over the gloo backend on CPU (the collectives on the
device are NCCL inside libvrt.so; this covers what the host decides before it hands over)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_grid_and_ranges_cover_everything():
    import bench
    for world in (1, 2, 4, 8, 16, 40):
        D, G = bench.shard_grid(world, 20)
        assert D * G == world and D <= 20
        lam = [bench.shard_range(91, G, g) for g in range(G)]
        assert lam[0][0] == 0 and lam[-1][1] == 91 and all(a[1] == b[0] for a, b in zip(lam, lam[1:]))
        assert max(b - a for a, b in lam) - min(b - a for a, b in lam) <= 1
        seen = np.concatenate([bench.round_robin(20, D, d) for d in range(D)])
        assert sorted(seen.tolist()) == list(range(20))


def test_lpt_balances_better_than_round_robin():
    import bench
    rng = np.random.default_rng(3)
    costs = list(rng.uniform(0.7, 1.4, 20))
    for D in (2, 4, 8):
        shards, load = bench.lpt_assign(costs, D)
        assert sorted(np.concatenate(shards).tolist()) == list(range(20))
        assert all(np.all(np.diff(s) > 0) for s in shards if len(s) > 1)        # each shard keeps the quadrature order
        rr = [sum(costs[i] for i in bench.round_robin(20, D, r)) for r in range(D)]
        assert max(load) <= max(rr) + 1e-12
        assert max(load) <= sum(costs) / D + max(costs)                         # the LPT guarantee


def test_split_assign_covers_every_direction_and_wavelength_once():
    import bench
    rng = np.random.default_rng(4)
    costs = list(rng.uniform(0.8, 1.5, 20))
    assert bench.split_assign(costs, 4, 91) is None and bench.split_assign(costs, 2, 91) is None      # 20 divides evenly
    assert bench.split_assign(costs, 3, 91) is None                                                   # 2 leftovers on 3 shards
    plan, load = bench.split_assign(costs, 8, 91)
    cover = np.zeros((20, 91), dtype=int)
    for whole, pieces in plan:
        assert len(whole) == 2 and len(pieces) == 1
        for i in whole:
            cover[i] += 1
        for i, lo, hi in pieces:
            assert hi - lo >= 16
            cover[i, lo:hi] += 1
    assert (cover == 1).all()
    _, lpt = bench.lpt_assign(costs, 8)
    assert max(load) / np.mean(load) < max(lpt) / np.mean(lpt)


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nq = 20
    D, G = bench.shard_grid(world, nq)
    di = rank % D
    mine = bench.round_robin(nq, D, di)
    true_cost = np.linspace(1.0, 2.0, nq) ** 2
    cost = torch.zeros(nq, dtype=torch.float64)
    cost[torch.as_tensor(mine)] = torch.as_tensor(true_cost[mine])       # every rank only knows the directions it built
    dist.all_reduce(cost, op=dist.ReduceOp.MAX)
    shards, load = bench.lpt_assign(list(cost.numpy()), D)
    got = torch.zeros(nq, dtype=torch.int64)
    got[torch.as_tensor(shards[di])] = 1
    dist.all_reduce(got)                                                    # every direction owned exactly once
    ok = bool((got == 1).all()) and np.allclose(cost.numpy(), true_cost)
    # the cell slices of the post-J stages tile the cells
    n = 1001
    cs = (n + D - 1) // D
    lo, hi = min(n, cs * di), min(n, cs * di + cs)
    cnt = torch.tensor([hi - lo]); dist.all_reduce(cnt)
    ok = ok and int(cnt) == n
    print("RANK", rank, "OK" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
''')


def test_two_rank_gloo_agreement(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", str(w), ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2, r.stdout
