#!/usr/bin/env python
"""Drop-in entry point with the reference's command line (model-single-GPU.py:736-741):

    python model-single-GPU.py --path . --batch_size 6

Same inputs (<path>/data/{dictionary.txt, offline-train.pkl, train_caption.txt, offline-test.pkl,
test_caption.txt}), same outputs (<path>/log-bs-6.txt, <path>/result/valid_decode_result-bs-6.txt,
<path>/result/valid-bs-6.wer), same schedule (Adadelta lr 1.0, validation every epoch with beam 10 /
maxlen 100, patience 15, lr/10 twice, then early stop; S:609-734) — but every FLOP runs in
libwap_b200.so on the B200.  Under torchrun (WORLD_SIZE > 1) it trains data-parallel with a NCCL
gradient all-reduce (semantics of model-distributed.py:704-717); rank 0 validates.
Extra flags (not in the reference): --max_epoch, --seed, --greedy_valid, --precision, --no_cuda_graphs, --prefetch N (default 2; 0 = synchronous
host prepare_data as in the reference), --save CKPT.npz (written by rank 0 whenever the validation WER improves - the reference's tf.train.Saver
is constructed but never called), --resume CKPT.npz (restores weights, Adadelta slots AND the lr / patience schedule).
CUDA-graph replay targets repeated batch shapes: a shape is captured the second time it is seen (fixed-shape workloads replay every step;
on the reference's area-bucketed batches most steps run as plain stream launches).
"""
import argparse
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(args):
    import torch
    import wap_b200
    from wap_b200 import dp, scorer
    rank, world, local = dp.init_distributed()
    torch.cuda.set_device(local)
    worddicts = wap_b200.load_dict(args.path + '/data/dictionary.txt')
    worddicts_r = [None] * len(worddicts)
    for kk, vv in worddicts.items():
        worddicts_r[vv] = kk
    train, train_uid_list = wap_b200.dataIterator(args.path + '/data/offline-train.pkl', args.path + '/data/train_caption.txt', worddicts,
                                                  batch_size=args.batch_size, batch_Imagesize=400000, maxlen=100, maxImagesize=400000)
    valid, valid_uid_list = wap_b200.dataIterator(args.path + '/data/offline-test.pkl', args.path + '/data/test_caption.txt', worddicts,
                                                  batch_size=args.batch_size, batch_Imagesize=400000, maxlen=100, maxImagesize=400000)
    print('train lenght is ', len(train))
    eng = wap_b200.WAPEngine(seed=args.seed, precision=args.precision, use_graphs=not args.no_cuda_graphs)
    eng.world_size, eng.rank = world, rank                 # rank enters the dropout seed: replicas draw different masks
    sched = eng.load(args.resume) if args.resume else {}
    dp.broadcast_params(eng.P.flat); dp.broadcast_params(eng.P.moving)
    if args.resume and world > 1:
        dp.broadcast_params(eng.accum); dp.broadcast_params(eng.accum_update)

    train = [b for b in train if b[0]]                       # drop empty batches BEFORE sharding: every rank runs the same number of steps
    # workspaces sized once for the largest bucket (growth - a multi-GB cudaMalloc - would otherwise land inside training steps)
    dims = [(len(bx), max(im.shape[1] for im in bx), max(im.shape[2] for im in bx), max(len(c) for c in by) + 1) for bx, by in train]
    if dims:
        for d in (max(dims, key=lambda d: d[0] * d[1] * d[2] * d[3]), max(dims, key=lambda d: d[0] * d[1] * d[2])):
            eng.reserve(*d)
    uidx, cost_s, dispFreq = 0, 0.0, 100
    validFreq = sampleFreq = max(len(train) // world, 1)
    history_errs, estop, halfLrFlag, patience, lrate, bad_counter = [], False, 0, 15, 1.0, 0
    first_epoch = 0
    if sched:                                                # --resume continues the schedule instead of restarting it at lr 1.0
        lrate, halfLrFlag, bad_counter = float(sched.get("lrate", lrate)), int(sched.get("halfLrFlag", 0)), int(sched.get("bad_counter", 0))
        history_errs = [float(v) for v in np.atleast_1d(sched.get("history_errs", []))]
        first_epoch, uidx = int(sched.get("epoch", 0)), int(sched.get("uidx", 0))
    log = open(args.path + '/log-bs-6.txt', 'w') if rank == 0 else None
    os.makedirs(args.path + '/result', exist_ok=True)
    rng = random.Random(args.seed)
    for _ in range(first_epoch):                             # resume: replay the shuffles of the finished epochs
        rng.shuffle(train)

    def say(msg):
        if rank == 0:
            print(msg); log.write(msg + '\n'); log.flush()

    bad = torch.zeros(1, device=eng.device)
    for epoch in range(first_epoch, args.max_epoch):
        rng.shuffle(train)                                   # same permutation on every rank; shards are disjoint
        shard = dp.shard_batches(train, rank, world)
        # --prefetch N (default 2): the next N batches are packed (raw uint8 pixels + captions, one pinned buffer each) on a background
        # thread; /255, padding and masks happen on the GPU (wap_prepare_batch).  --prefetch 0: the reference's synchronous host path.
        prepared = wap_b200.Prefetcher(shard, pinned=True, depth=args.prefetch, packed=True) if args.prefetch > 0 else \
            (wap_b200.prepare_data(bx, by, pinned=True) for bx, by in shard)
        for item in prepared:
            uidx += 1
            if args.prefetch > 0:
                cost_i = float(eng.train_step_packed(item, lr=lrate))
            else:
                x, x_m, y, y_m = item
                cost_i = float(eng.train_step(x, y, x_m, y_m, lr=lrate))
            cost_s += cost_i
            bad[0] = float(np.isnan(cost_i) or np.isinf(cost_i))
            if world > 1:                                    # every rank must leave together, or the others hang in the next all-reduce
                torch.distributed.all_reduce(bad, op=torch.distributed.ReduceOp.MAX)
            if float(bad[0]) > 0:
                print('invalid cost value detected')
                if world > 1:
                    torch.distributed.destroy_process_group()
                sys.exit(0)
            if uidx % dispFreq == 0:
                cost_s /= dispFreq
                say('Epoch ' + str(epoch) + ' Update ' + str(uidx) + ' Cost ' + str(cost_s) + ' Lr ' + str(lrate))
                cost_s = 0.0
            if uidx % sampleFreq == 0:
                dp.sync_moving_stats(eng.P.moving)
            if uidx % sampleFreq == 0 and rank == 0:
                with open(args.path + '/result/valid_decode_result-bs-6.txt', 'w') as fpp, eng.decoding():
                    idx = 0
                    def emit(ss):
                        nonlocal idx
                        fpp.write(valid_uid_list[idx]); idx += 1
                        if idx % 100 == 0:
                            say('gen %d samples' % idx)
                        for vv in ss:
                            if vv == 0:
                                break
                            fpp.write(' ' + worddicts_r[vv])
                        fpp.write('\n')
                    # one unpadded image at a time as in the reference (S:661-672), as a two-stream pipeline: the encoder of image i+1 runs next to
                    # the search / token loop of image i; results are read from the device once per validation pass
                    imgs = [torch.as_tensor((np.moveaxis(xx, 0, -1) / 255.).astype(np.float32)[None]).to(eng.device) for vx, _ in valid for xx in vx]
                    if args.greedy_valid:
                        for toks, lens in eng.greedy_decode_images(imgs, 100):
                            emit(toks[0, :int(lens[0])].tolist())
                    else:
                        for ss in eng.get_sample_images(imgs, 10, 100):
                            emit(ss)
                say('valid set decode done')
            if uidx % validFreq == 0 and rank == 0:
                probs = []
                for vx, vy in valid:
                    if not vx:
                        continue
                    x, x_m, y, y_m = wap_b200.prepare_data(vx, vy)
                    ann, a_m = eng.dense_net(x, x_m, training=False)
                    l2 = 1e-4 * float((eng.P.flat[:eng.P.n_decay].double() ** 2).sum())     # the reference's cost includes L2 (S:592-598)
                    probs.append(float(eng.get_cost(ann, y, a_m, y_m, dropout=False)) + l2)
                valid_err_cost = float(np.mean(probs))
                wer, sacc = scorer.process(args.path + '/result/valid_decode_result-bs-6.txt', args.path + '/data/test_caption.txt',
                                           args.path + '/result/valid-bs-6.wer')
                valid_per, valid_sacc = 100. * wer, 100. * sacc
                history_errs.append(valid_per)
                if valid_per <= np.array(history_errs).min():
                    bad_counter = 0
                    if args.save:
                        eng.save(args.save, schedule=dict(lrate=lrate, halfLrFlag=halfLrFlag, bad_counter=bad_counter, history_errs=history_errs,
                                                          epoch=epoch + 1, uidx=uidx))
                else:
                    bad_counter += 1
                    if bad_counter > patience:
                        if halfLrFlag == 2:
                            say('Early Stop!'); estop = True
                        else:
                            say('Lr decay and retrain!'); bad_counter = 0; lrate = lrate / 10; halfLrFlag += 1
                say('Valid WER: %.2f%%, ExpRate: %.2f%%, Cost: %f' % (valid_per, valid_sacc, valid_err_cost))
            if world > 1 and uidx % validFreq == 0:          # schedule decisions are taken on rank 0
                flags = torch.tensor([lrate, float(estop)], device=eng.device, dtype=torch.float64)
                torch.distributed.broadcast(flags, 0)
                lrate, estop = float(flags[0]), bool(flags[1] > 0)
            if estop:
                break
        if estop:
            break


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--path", default=".")
    parser.add_argument("--batch_size", type=int, default=6)
    parser.add_argument("--max_epoch", type=int, default=200)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--greedy_valid", action="store_true")
    parser.add_argument("--save", default="")
    parser.add_argument("--prefetch", type=int, default=2, help="pack this many batches ahead on a background thread (GPU-side /255 + masks); 0 = host prepare_data")
    parser.add_argument("--no_cuda_graphs", action="store_true", help="never capture/replay forward+backward (default: shapes seen twice are captured)")
    parser.add_argument("--cuda_graphs", action="store_true", help="(kept for round-1 command lines; graph replay is the default now)")
    parser.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32", "tf32x3", "bf16"])
    parser.add_argument("--resume", default="")
    (args, unknown) = parser.parse_known_args()
    main(args)
