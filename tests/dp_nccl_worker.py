"""torchrun worker of tests/test_gpu_cfg2.py::test_dp_two_ranks_nccl (2 ranks, NCCL).

Each rank runs WAPEngine.train_step on ITS shard with world_size = 2 (all-reduce(sum) of the flat gradient + the 1/2 mean
folded into the optimiser kernel = SyncReplicasOptimizer's mean of replica gradients, model-distributed.py:704-717).  Checks:
  * both ranks end with bit-identical weights and Adadelta state;
  * those weights equal a single-process update that is handed the SUM of the two shard gradients (computed on this rank
    from both shards, same seeds), i.e. the collective changed nothing but where the gradients came from."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import wap_b200
    from helpers import O, make_batch
    from wap_b200 import dp
    rank, world, local = dp.init_distributed("nccl")
    assert world == 2
    torch.cuda.set_device(local)
    P = O.lively_params(2)
    shards = [make_batch(21, 2, 40, 72, 4, True), make_batch(22, 2, 40, 72, 4, True)]
    seeds = [71, 72]
    eng = wap_b200.WAPEngine(init=False); eng.P.load(P)
    eng.world_size, eng.rank = world, rank
    x, xm, y, ym = shards[rank]
    cost = float(eng.train_step(x, y, xm, ym, lr=1.0, dropout=True, seed=seeds[rank]))
    # --- identical replicas
    mine = torch.cat([eng.P.flat, eng.accum, eng.accum_update])
    other = mine.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(mine, other))
    # --- single-process statement of the same update
    ref = wap_b200.WAPEngine(init=False); ref.P.load(P); ref.world_size = 2
    ref.forward_backward(shards[0][0], shards[0][2], shards[0][1], shards[0][3], dropout=True, seed=seeds[0]); g0 = ref.grads.clone()
    ref.P.load(P)
    ref.forward_backward(shards[1][0], shards[1][2], shards[1][1], shards[1][3], dropout=True, seed=seeds[1])
    ref.grads += g0
    ref.apply_update(1.0)
    # two fp32 runs of the same backward differ by the order of their atomic adds (~1e-6 relative); Adadelta's first step
    # (eps = 1e-8) turns that into a small absolute difference of the update, so compare on the update scale (lr * O(1e-3))
    dp_ = float((eng.P.flat - ref.P.flat).abs().max())
    scale = float((ref.P.flat - _flat_of(ref, P)).abs().max())
    print(f"rank {rank}: cost {cost:.6f} identical_replicas={same} max|dp - single| = {dp_:.3e} (update scale {scale:.3e})", flush=True)
    ok = same and np.isfinite(cost) and dp_ <= 2e-2 * scale
    dist.barrier()
    if ok:
        print(f"DP_NCCL_OK {rank}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def _flat_of(eng, P):
    import wap_b200
    tmp = wap_b200.engine.ParamStore(eng.device)
    tmp.load(P)
    return tmp.flat


if __name__ == "__main__":
    main()
