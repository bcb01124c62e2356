"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: k-slice partition + all-gather of the
S(k) slices reproduces the single-process structure factor; replicas get distinct seeds."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import replay
from plum_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_py import Oracle
    r, s, types, params = replay.load_golden(name)
    o = Oracle(params, repl_mode=1)
    o.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    full_ref = o.sk_half()
    n_k = full_ref.shape[0]

    def compute_slice(first, count):
        return full_ref[first:first + count]   # a rank only ever touches its own slice

    sk = sharded.gather_sk(compute_slice, n_k, rank, world)
    e_local = torch.tensor([float(np.sum(full_ref[slice(*_rng(sharded.k_slice(n_k, rank, world)))] ** 2))],
                           dtype=torch.float64)
    dist.all_reduce(e_local)
    np.save(os.path.join(out_dir, f"sk_{rank}.npy"), sk.numpy())
    np.save(os.path.join(out_dir, f"e_{rank}.npy"), np.array([e_local.item(), float(np.sum(full_ref ** 2))]))
    dist.destroy_process_group()


def _rng(fc):
    return fc[0], fc[0] + fc[1]


def test_k_slices_partition_the_list():
    for n_k in (0, 1, 7, 13, 46, 1787, 2069):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                f, c = sharded.k_slice(n_k, r, world)
                seen += list(range(f, f + c))
                assert c <= sharded.padded_count(n_k, world)
            assert seen == list(range(n_k))


def test_replica_seeds_distinct():
    assert len({sharded.replica_seed(1, r) for r in range(8)}) == 8


@pytest.mark.parametrize("name", ["bulk_nvt", "confined_nvt"])
def test_world2_allgather_reproduces_full_structure_factor(name, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    from oracle.oracle_py import Oracle
    r, s, types, params = replay.load_golden(name)
    o = Oracle(params, repl_mode=1)
    o.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    ref = o.sk_half()
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), f"sk_{rank}.npy"))
        assert np.array_equal(got, ref)          # slices are moved, never recomputed: bit identical
        e = np.load(os.path.join(str(tmp_path), f"e_{rank}.npy"))
        assert abs(e[0] - e[1]) <= 1e-12 * abs(e[1])


def _attach_worker(rank, world, port, out):
    import torch.distributed as dist
    from plum_b200 import sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class FakeEngine:
        """Stands in for Engine: records what attach_peers hands to pg_sk_attach."""
        def sk_export(self):
            return bytes([rank]) * 64

        def sk_attach(self, rk, w, handles):
            out.put((rank, rk, w, handles))

    sharded.attach_peers(FakeEngine(), rank, world)
    dist.destroy_process_group()


def test_attach_peers_exchanges_every_ranks_handle_in_rank_order():
    """Host plumbing of the fused k-sharded recompute (world_size 2, gloo): every rank receives all 64-byte handles,
    concatenated in rank order, together with its own rank and the world size."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29731
    ps = [ctx.Process(target=_attach_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes([0]) * 64 + bytes([1]) * 64
    assert got == [(0, 0, 2, want), (1, 1, 2, want)]
