"""world_size-2 gloo test of the data-parallel host logic: sharding the step's positives over
ranks + one all-reduce of the table gradient reproduces the single-process step; every rank then applies
the identical RSGD update.  The per-rank arithmetic is the CPU oracle (there is no GPU here); on the GPU
box the same functions run over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from learning_embeddings_b200 import sharding
from oracle import cones


def test_shard_bounds_partition_everything():
    for n in (0, 1, 7, 8, 1974, 190650):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = load_golden("step_hyp_D10_a0p05")
    Nn, K, alpha = int(g["N"]), float(g["K"]), float(g["alpha"])
    B = 600
    u, v = g["u"][:B], g["v"][:B]
    drawn = g["drawn"].reshape(-1, Nn, 2)[:B]
    neg_to, neg_from = drawn[:, :, 0], drawn[:, :, 1]
    su, sv, snt, snf = sharding.shard_groups(u, v, neg_to, neg_from, rank, world)
    nf = np.concatenate([np.repeat(su[:, None], Nn, 1), snf], 1).reshape(-1)
    nt = np.concatenate([snt, np.repeat(sv[:, None], Nn, 1)], 1).reshape(-1)
    W = torch.from_numpy(g["W0"]).double()
    r = cones.label_step("hyp", W, cones.ROW_HYP_SHELL, K, alpha, torch.from_numpy(su), torch.from_numpy(sv),
                         torch.from_numpy(nf), torch.from_numpy(nt))
    grad, loss = r["gW"].clone(), r["loss"].reshape(1).clone()
    sharding.allreduce_grad_and_loss(grad, loss)
    _, W_new = cones.rsgd_step(W, grad, 0.01, cones.inner_radius(K))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), grad=grad.numpy(), loss=loss.numpy(), W_new=W_new.numpy())
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = load_golden("step_hyp_D10_a0p05")
    Nn, K, alpha = int(g["N"]), float(g["K"]), float(g["alpha"])
    B = 600
    u, v = g["u"][:B], g["v"][:B]
    drawn = g["drawn"].reshape(-1, Nn, 2)[:B]
    nf = np.concatenate([np.repeat(u[:, None], Nn, 1), drawn[:, :, 1]], 1).reshape(-1)
    nt = np.concatenate([drawn[:, :, 0], np.repeat(v[:, None], Nn, 1)], 1).reshape(-1)
    W = torch.from_numpy(g["W0"]).double()
    ref = cones.label_step("hyp", W, cones.ROW_HYP_SHELL, K, alpha, torch.from_numpy(u), torch.from_numpy(v),
                           torch.from_numpy(nf), torch.from_numpy(nt))
    _, W_ref = cones.rsgd_step(W, ref["gW"], 0.01, cones.inner_radius(K))
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for o in outs:
        np.testing.assert_allclose(o["grad"], ref["gW"].numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(float(o["loss"][0]), float(ref["loss"]), rtol=1e-12)
        np.testing.assert_allclose(o["W_new"], W_ref.numpy(), rtol=1e-12, atol=1e-14)
    assert np.array_equal(outs[0]["W_new"], outs[1]["W_new"])  # replicas stay identical without a broadcast
