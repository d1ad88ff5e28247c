"""Multi-rank path on CPU (gloo, world_size 2): rank 0 generates the evaluation keys, the other rank only allocates
them (`set_keys_external`) and receives them by broadcast; both ranks then evaluate the same ciphertext arithmetic on
their own shard and must agree bit for bit with a single-rank engine (independent ciphertext pairs per
rank, one key broadcast at setup, no data-path collective)."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, q):
    sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200"), str(ROOT / "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    import backend
    mod = backend.use_emulation()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = mod.Engine(logn=12, levels=5, dnum=2, hamming_weight=64, seed=21)
    eng.set_keys_external(rank != 0)
    sk = eng.create_secret_key()
    eng.create_public_key(sk)
    rk = eng.create_relinearization_key(sk)
    eng.create_conjugation_key(sk)
    eng.create_rotation_key(sk, [3])
    moved = eng.broadcast_evaluation_keys(dist, src=0)
    rng = np.random.default_rng(100 + rank)                         # each rank has its own shard
    z = np.exp(2j * np.pi * rng.random(eng.slot_count))
    ct = eng.encrypt(z)
    out = eng.rotate(eng.conjugate(eng.multiply(ct, ct, rk)), None, 3)
    err = float(np.abs(eng.decrypt(out) - np.roll(np.conj(z * z), 3)).max())
    a = np.zeros((2, out.level + 1, 4096), dtype=np.uint64)
    eng._lib.ckks_ct_export(eng._ptr, out._h, a)
    q.put((rank, moved, err, int(a.sum(dtype=np.uint64))))
    dist.barrier()
    dist.destroy_process_group()


def test_key_broadcast_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] > 0                     # same number of key bytes on both ranks
    assert res[0][2] < 1e-7 and res[1][2] < 1e-7          # rank 1 evaluates correctly with the broadcast keys
    # reference: a single engine that generated its own keys, evaluating rank 1's shard
    sys.path[:0] = [str(ROOT / "tests")]
    import backend
    mod = backend.use_emulation()
    eng = mod.Engine(logn=12, levels=5, dnum=2, hamming_weight=64, seed=21)
    sk = eng.create_secret_key(); eng.create_public_key(sk); rk = eng.create_relinearization_key(sk)
    rng = np.random.default_rng(101)
    z = np.exp(2j * np.pi * rng.random(eng.slot_count))
    ct = eng.encrypt(z)
    out = eng.rotate(eng.conjugate(eng.multiply(ct, ct, rk)), None, 3)
    a = np.zeros((2, out.level + 1, 4096), dtype=np.uint64)
    eng._lib.ckks_ct_export(eng._ptr, out._h, a)
    assert int(a.sum(dtype=np.uint64)) == res[1][3]       # bit-identical ciphertext: the broadcast keys are the same keys
