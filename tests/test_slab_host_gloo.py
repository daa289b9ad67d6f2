"""Host-side logic of the multi-GPU path on CPU: two ranks over gloo agree on the
slab partition, bootstrap a 128-byte id through torch.distributed exactly like
bench.py does, and generate disjoint slabs whose union is the global scene."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # the id bootstrap of bench.py (there the payload is ncclGetUniqueId's 128 bytes)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(idt, 0)
    import smoothed_particle_hydrodynamics_b200 as S
    sc = bench.column_scene(S, world, rank, True, sites_xy=(32, 16), planes=64)
    # the integrity record of bench.py's slab lines: count / sum / xor of the owned ids over all ranks
    cnt, idsum, idxor = bench.id_checksums(sc["gids"])
    red = torch.tensor([cnt, idsum], dtype=torch.int64)
    dist.all_reduce(red)
    xors = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(xors, torch.tensor([idxor], dtype=torch.int64))
    x = 0
    for t in xors:
        x ^= int(t.item())
    ok = bench.id_checksums_ok(int(red[0]), int(red[1]), x, sc["total"])
    assert ok["owned_total"] == ok["expected"] and ok["id_sum_ok"] and ok["id_xor_ok"], ok
    bad = bench.id_checksums_ok(int(red[0]), int(red[1]) + 1, x, sc["total"])       # one id off: caught
    assert not bad["id_sum_ok"]
    n_local = torch.tensor([sc["gids"].size], dtype=torch.int64)
    dist.all_reduce(n_local)
    gz = torch.tensor([sc["grid"][2]], dtype=torch.int64)
    dist.all_reduce(gz, op=dist.ReduceOp.MAX)
    q.put((rank, idt.numpy().tobytes(), sc["layers"], sc["grid"], sc["gids"], sc["pos"], int(n_local), int(gz),
           sc["total"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_and_bootstrap():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, id0, layers0, grid0, g0, p0, n0, gz0, tot0), (r1, id1, layers1, grid1, g1, p1, n1, gz1, tot1) = out
    assert id0 == id1 == bytes(range(128))
    assert layers0 == layers1 and grid0 == grid1 and gz0 == grid0[2]
    assert layers0[0][0] == 0 and layers0[-1][1] == grid0[2] and layers0[0][1] == layers0[1][0]
    assert n0 == n1 == tot0 == g0.size + g1.size                      # every particle owned exactly once
    assert np.intersect1d(g0, g1).size == 0
    # each rank's particles are the global scene's particles with those ids
    from oracle import scenes
    d = scenes.lattice_spacing(0.1, 40.0)
    glob = scenes.lattice_scene(32, 16, 64, d, origin=(np.float32(49 * 0.2), np.float32(16 * 0.2), 3 * 0.2))
    assert np.array_equal(glob[g0], p0) and np.array_equal(glob[g1], p1)
    # and they lie in the voxel layers the rank owns
    inv2h = np.float32(1.0) / (np.float32(0.1) * np.float32(2.0))
    for (z0, z1), p in zip(layers0, (p0, p1)):
        vz = np.clip(np.floor(p[:, 2] * inv2h), 0, grid0[2] - 1)
        assert vz.min() >= z0 and vz.max() < z1
