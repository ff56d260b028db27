"""Run under torchrun on N GPUs: the sharded step (pairs split over ranks + NCCL all-reduce of the table
gradient) must give the same table as the same batch processed by one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multigpu_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_embeddings_b200 import hierarchy as H, sharding  # noqa: E402
from learning_embeddings_b200.engine import ConeStep  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h = H.ethec()
    edges = h.closure_edges()
    rng = np.random.default_rng(0)  # same batch on every rank
    B, Nn = 40000, 5
    sel = rng.integers(0, len(edges), size=B)
    u, v = edges[sel, 0], edges[sel, 1]
    nt, nf = h.sample_negatives(u, v, Nn, rng)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    ok = True
    # "p2p" = the packet exchange inside the fused update kernel (lec_update_rows); "nccl" = all_reduce between the kernels
    cases = [("hyp", 10, "rsgd", 1, "nccl"), ("hyp", 10, "rsgd", 3, "nccl"), ("hyp", 10, "rsgd", 1, "p2p"),
             ("hyp", 10, "rsgd", 4, "p2p"), ("hyp", 10, "rsgd", 5, "p2p"), ("euc", 2, "adam", 4, "p2p"),
             ("euc", 10, "sgd", 3, "p2p"), ("hyp", 50, "rsgd", 3, "p2p"), ("hyp", 10, "rsgd", 4, "two_shot"),
             ("hyp", 50, "rsgd", 3, "two_shot"), ("euc", 10, "adam", 3, "two_shot")]
    for geom, D, update, step_count, comm in cases:
        g = torch.Generator().manual_seed(0)
        w = torch.randn(h.n, D, generator=g)
        K = 0.1 if geom == "hyp" else 3.0
        W0 = (0.0990195 + 0.05 * torch.rand(h.n, 1, generator=g)) * w / w.norm(dim=1, keepdim=True) if geom == "hyp" else w
        # sharded
        Ws = W0.to(dev).clone()
        eng = ConeStep(Ws, geom, Nn, B, K=K, alpha=0.05, lr=0.01, update=update, process_group=dist.group.WORLD,
                       comm="p2p" if comm == "two_shot" else comm, exchange_mode=sharding.TWO_SHOT if comm == "two_shot" else None)
        if rank == 0:
            print("%s D=%d %s: comm requested %s -> %s %s" % (geom, D, update, comm, eng.comm, eng.comm_note), flush=True)
        su, sv, snt, snf = sharding.shard_groups(u, v, nt, nf, rank, world)
        for _ in range(step_count):
            eng.step_device(to_dev(su), to_dev(sv), to_dev(snt), to_dev(snf))
        loss_s = float(eng.global_loss().item())
        # single GPU, whole batch
        W1 = W0.to(dev).clone()
        one = ConeStep(W1, geom, Nn, B, K=K, alpha=0.05, lr=0.01, update=update)
        for _ in range(step_count):
            one.step_device(to_dev(u), to_dev(v), to_dev(nt), to_dev(nf))
        loss_1 = float(one.loss.item())
        torch.cuda.synchronize()
        err = float((Ws - W1).abs().max())
        rel = abs(loss_s - loss_1) / abs(loss_1)
        # replicas identical across ranks
        ref = Ws.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(ref, Ws))
        print("rank %d/%d %s D=%d %s steps=%d  max|W_sharded - W_single| = %.3e  loss rel diff %.2e  identical across ranks: %s"
              % (rank, world, geom, D, update, step_count, err, rel, same), flush=True)
        # fp32 L2 reductions are order-dependent: two runs of the SAME single-GPU step differ by ~7e-6 on this
        # batch (gradient sums ~1e3 with heavy cancellation, x lr/lambda^2), so that is the resolution here
        if update == "adam":
            # Adam moves an element by ~lr * sign(g) whatever |g| is: the rare element whose gradient cancels to rounding
            # noise may land elsewhere when the summation order changes -- bound the bulk, count the outliers
            far = float(((Ws - W1).abs() > 1e-5).float().mean())
            ok = ok and far < 0.005 and rel < 1e-6 and same
        else:
            # plain SGD on Euclidean cones moves a row by lr * (a gradient sum of ~1e4 in magnitude): its fp32 summation
            # noise is ~1e-4 where RSGD's (rescaled by (1-|w|)^2/4) is ~1e-5
            ok = ok and err < (5e-4 if update == "sgd" else 5e-5) and rel < 1e-6 and same
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("multi-GPU parity FAILED")
    if rank == 0:
        print("multi-GPU parity ok")


main()
