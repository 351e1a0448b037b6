"""Two real GPUs, NCCL: sharded evaluation equals the single-GPU result (skipped on a 1-GPU box)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(kind):
    from oracle import workloads as wl

    if kind == "syn130":  # two-stage tridiagonalisation: band arrays collected over several chunks before the peer push
        return wl.synthetic(130, 3, seed=9)
    return wl.haldane() if kind == "haldane" else wl.synthetic(12, 10)


def _worker(rank, world, port, n_k, out_dir, kind="syn12"):
    import sys

    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import workloads as wl
    from tbmodels_b200.sharded import ShardedEvaluator, broadcast_model

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        packed = broadcast_model(_model(kind) if rank == 0 else None, src=0, device=f"cuda:{rank}")
        sh = ShardedEvaluator(packed, device=rank)
        g = torch.Generator(device=f"cuda:{rank}").manual_seed(7)  # same k-points on every rank
        k_all = torch.rand((n_k, packed.dim), dtype=torch.float64, device=f"cuda:{rank}", generator=g)
        lo, hi, e_loc = sh.eigenval_local(k_all)
        full = sh.eigenval_allgather(k_all)
        fused = sh.eigenval_allgather_fused(k_all).clone()   # peer stores over NVLink + device-side barrier
        fused2 = sh.eigenval_allgather_fused(k_all).clone()  # buffer reuse
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, hi=hi, e_loc=e_loc.cpu().numpy(), full=full.cpu().numpy(),
                 k=k_all.cpu().numpy(), fused=fused.cpu().numpy(), fused2=fused2.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_k,ws_mb,kind", [(4096, None, "syn12"), (5001, None, "syn12"), (30001, "1", "syn12"),
                                           (2_500_001, None, "haldane"), (301, "1", "syn130")])
def test_two_gpu_nccl_sharding(tmp_path, monkeypatch, n_k, ws_mb, kind):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import tbmodels_b200 as tbk
    from oracle import workloads as wl

    if ws_mb:  # tiny workspace: many chunks per rank, so the peer stores of the fused gather interleave with compute
        monkeypatch.setenv("TBK_WORKSPACE_MB", ws_mb)
    if kind == "syn130":  # groups of four chunks, several groups per rank
        monkeypatch.setenv("TBK_BAND_GROUP_MB", "1")
    mp.spawn(_worker, args=(2, _free_port(), n_k, str(tmp_path), kind), nprocs=2, join=True)
    monkeypatch.delenv("TBK_WORKSPACE_MB", raising=False)
    monkeypatch.delenv("TBK_BAND_GROUP_MB", raising=False)
    d0 = np.load(tmp_path / "r0.npz")
    d1 = np.load(tmp_path / "r1.npz")
    want = tbk.Evaluator(_model(kind), device=0).eigenval_array(d0["k"])
    assert np.array_equal(d0["k"], d1["k"])
    for d in (d0, d1):
        assert np.array_equal(d["full"], want)  # bit-identical regardless of the shard a k-point landed in
        assert np.array_equal(d["fused"], want) and np.array_equal(d["fused2"], want)  # fused peer-store gather too
        assert np.array_equal(d["e_loc"], want[int(d["lo"]) : int(d["hi"])])
    assert int(d0["hi"]) == int(d1["lo"]) and int(d1["hi"]) == n_k
