"""N > 1 host logic on CPU: world_size-2 gloo groups, the oracle standing in for the per-rank GPU evaluator."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds_cover_batch_in_order():
    from tbmodels_b200.sharded import max_shard, shard_bounds

    for n_k in (0, 1, 7, 8, 9, 1000, 10**8 + 3):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(n_k, world, r)
                assert lo == prev and hi >= lo and hi - lo <= max_shard(n_k, world)
                prev = hi
            assert prev == n_k
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


MESH = (5, 7)  # 5 lines of 7 points: ranks get 3 + 2 lines


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_k, result_dir):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from oracle import tb_oracle as orc
    from oracle import workloads as wl
    from tbmodels_b200.sharded import ShardedEvaluator, broadcast_model

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        packed = broadcast_model(wl.haldane() if rank == 0 else None, src=0)

        class OracleLocal:  # stand-in for the CUDA Evaluator: same method contract, CPU tensors
            def eigenval_device(self, k):
                return torch.from_numpy(orc.eigenval_array(packed.R, packed.hop, packed.pos, k.numpy()))

            def hamilton_device(self, k, convention=2):
                return torch.from_numpy(orc.hamilton(packed.R, packed.hop, packed.pos, k.numpy(), convention))

            def eigenval_mesh_device(self, dims, shift=None, first_line=0, n_lines=None):
                k = wl.kgrid_points(dims, shift)[first_line * dims[-1]:(first_line + n_lines) * dims[-1]]
                return torch.from_numpy(orc.eigenval_array(packed.R, packed.hop, packed.pos, k))

        sh = ShardedEvaluator(packed, local=OracleLocal())
        k_all = torch.from_numpy(np.random.default_rng(0).random((n_k, 2)))
        lo, hi, e_loc = sh.eigenval_local(k_all)
        full = sh.eigenval_allgather(k_all)
        lo2, hi2, h_loc = sh.hamilton_local(k_all, convention=1)
        mlo, mhi, e_mesh = sh.eigenval_mesh_local(MESH, shift=(0.0, 0.5))
        np.savez(
            os.path.join(result_dir, f"r{rank}.npz"), lo=lo, hi=hi, e_loc=e_loc.numpy(), full=full.numpy(),
            h_loc=h_loc.numpy(), R=packed.R, hop=packed.hop, mlo=mlo, mhi=mhi, e_mesh=e_mesh.numpy(),
        )
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_k", [64, 101])
def test_two_rank_gloo_sharding(tmp_path, n_k):
    import torch.multiprocessing as mp

    from oracle import tb_oracle as orc
    from oracle import workloads as wl

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_k, str(tmp_path)), nprocs=world, join=True)
    p = wl.haldane()
    k_all = np.random.default_rng(0).random((n_k, 2))
    want = orc.eigenval_array(p.R, p.hop, p.pos, k_all)
    want_h = orc.hamilton(p.R, p.hop, p.pos, k_all, 1)
    want_mesh = orc.eigenval_array(p.R, p.hop, p.pos, wl.kgrid_points(MESH, (0.0, 0.5)))
    mesh_prev = 0
    covered = 0
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        lo, hi = int(d["lo"]), int(d["hi"])
        assert np.array_equal(d["e_loc"], want[lo:hi])
        assert np.array_equal(d["h_loc"], want_h[lo:hi])
        assert np.array_equal(d["full"], want)  # every rank holds the whole, ordered result
        assert np.array_equal(d["hop"], p.hop) and np.array_equal(d["R"], p.R)  # model broadcast
        covered += hi - lo
        mlo, mhi = int(d["mlo"]), int(d["mhi"])  # whole lines, contiguous, in rank order
        assert mlo == mesh_prev and mlo % MESH[-1] == 0 and mhi % MESH[-1] == 0
        assert np.array_equal(d["e_mesh"], want_mesh[mlo:mhi])
        mesh_prev = mhi
    assert covered == n_k and mesh_prev == MESH[0] * MESH[1]
