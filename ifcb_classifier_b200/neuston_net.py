#!/usr/bin/env python
"""``neuston_net.py``-compatible command line (reference neuston_net.py:311-452).

  python -m ifcb_classifier_b200.neuston_net [--batch N] [--loaders N] RUN SRC MODEL RUN_ID
         [--type bin|img] [--outdir ...] [--outfile ...]... [--filter IN|OUT KW...] [--clobber] [--gobig]
  torchrun --nproc-per-node 8 -m ifcb_classifier_b200.neuston_net RUN ...   # bins sharded across GPUs

Same flags, defaults and path templates ({RUN_ID} {RUN_DATE} {MODEL_ID} {BIN_ID} {BIN_YEAR}
{BIN_DATE} {INPUT_SUBDIRS}).  The loop that pytorch_lightning's Trainer ran is
``engine.BinClassifier``; per-bin error isolation and skip-if-exists are preserved
(neuston_net.py:242-251,258-268).  TRAIN runs ``train_loop.do_training``: the reference's
dataset / split / epoch / checkpoint flow around the B200 train step (``train.TrainNet``);
``torchrun --nproc-per-node N ... TRAIN ...`` is data parallel (NCCL gradient mean).
"""
import argparse
import datetime as dt
import os
import sys
import time


def argparse_nn(parser=None):
    if parser is None:
        parser = argparse.ArgumentParser(description='Train and run IFCB image classifiers on B200 GPUs')
    sub = parser.add_subparsers(dest='cmd_mode', help='optional arguments below must precede TRAIN / RUN')
    train = sub.add_parser('TRAIN', help='Train a new model')
    run = sub.add_parser('RUN', help='Run a previously trained model')
    common = parser.add_argument_group(title='NN Common Args')
    common.add_argument('--batch', dest='batch_size', metavar='SIZE', default=108, type=int)
    common.add_argument('--loaders', metavar='N', default=4, type=int)
    common.add_argument('--dtype', default='fp16', choices=['fp16', 'bf16'],
                        help='tensor-core operand format of eval-mode forward passes (B200 extension; default fp16)')
    common.add_argument('--deterministic', action='store_true',
                        help='TRAIN: bitwise reproducible steps (ordered reductions instead of floating-point atomics; the reference '
                             'runs Trainer(deterministic=True)); a few %% slower (B200 extension)')
    common.add_argument('--train-dtype', dest='train_dtype', default='bf16', choices=['bf16'],
                        help='16-bit storage format of the TRAIN step (B200 extension)')
    _train_args(train)
    _run_args(run)
    return parser


def _train_args(p):
    p.add_argument('SRC'); p.add_argument('MODEL'); p.add_argument('TRAIN_ID')
    p.add_argument('--untrain', dest='pretrained', default=True, action='store_false')
    p.add_argument('--img-norm', nargs=2, metavar=('MEAN', 'STD'))
    p.add_argument('--seed', default=0, type=int)
    p.add_argument('--split', metavar='T:V', default='80:20')
    p.add_argument('--class-config', metavar=('CSV', 'COL'), nargs=2)
    p.add_argument('--class-min', metavar='MIN', default=2, type=int)
    p.add_argument('--class-max', metavar='MAX', default=None, type=int)
    p.add_argument('--swap', default=False, action='store_true', help=argparse.SUPPRESS)
    p.add_argument('--emax', metavar='MAX', default=60, type=int)
    p.add_argument('--emin', metavar='MIN', default=10, type=int)
    p.add_argument('--estop', metavar='STOP', default=10, type=int)
    p.add_argument('--flip', choices=['x', 'y', 'xy', 'x+V', 'y+V', 'xy+V'])
    p.add_argument('--outdir', default='training-output/{TRAIN_ID}')
    p.add_argument('--model-id', default='{TRAIN_ID}')
    p.add_argument('--epochs-log', metavar='ELOG', default='epochs.csv')
    p.add_argument('--args-log', metavar='ALOG', default='args.yml')
    p.add_argument('--onnx', action='store_true')
    p.add_argument('--results', dest='result_files', metavar=('FNAME', 'SERIES'), nargs='+', action='append')
    p.add_argument('--dataset-id'); p.add_argument('--notes')


def _run_args(p):
    p.add_argument('SRC'); p.add_argument('MODEL'); p.add_argument('RUN_ID')
    p.add_argument('--type', dest='src_type', default='bin', choices=['bin', 'img'])
    p.add_argument('--outdir', default='run-output/{RUN_ID}/v3/{MODEL_ID}')
    p.add_argument('--outfile', action='append')
    p.add_argument('--filter', nargs='+', metavar=('IN|OUT', 'KEYWORD'))
    p.add_argument('--clobber', action='store_true')
    p.add_argument('--gobig', action='store_true', help=argparse.SUPPRESS)


def argparse_nn_runtimeparams(args, classifier=None):
    args.cmd_timestamp = dt.datetime.now(dt.timezone.utc).isoformat(timespec='seconds')
    try:
        with open('version') as f:
            args.version = f.read().strip()
    except FileNotFoundError:
        args.version = None
    import torch
    if torch.cuda.is_available():
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        args.gpus = [int(g) for g in vis.split(',')] if vis else list(range(torch.cuda.device_count()))
    else:
        args.gpus = None
    date_str = args.cmd_timestamp.split('T')[0]
    if args.cmd_mode == 'TRAIN':
        args.outdir = args.outdir.format(TRAIN_DATE=date_str, TRAIN_ID=args.TRAIN_ID)
    elif args.cmd_mode == 'RUN':
        model_id = getattr(classifier.hparams, 'model_id', None) if classifier is not None else None
        args.outdir = args.outdir.format(RUN_DATE=date_str, RUN_ID=args.RUN_ID, MODEL_ID=model_id)


def _engine_batch(args, arch):
    """ROIs per network launch sequence.  The reference's ``--batch`` (default 108) sizes a DataLoader batch for a 2019
    GPU; a B200 needs ~1k ROIs in flight to fill 148 SMs, so the default is raised to 1024 -- results do not depend on it.
    An explicit ``--batch`` (or IFCB_RUN_BATCH) is honoured."""
    env = os.environ.get('IFCB_RUN_BATCH')
    if env:
        return max(int(env), 16)
    if args.batch_size != 108:
        return max(args.batch_size, 16)
    # one activation buffer per layer: keep the deep / wide families within a few tens of GB of HBM
    return {'inception_v3': 1024, 'resnet18': 1024, 'resnet34': 1024, 'resnet50': 1024, 'resnet101': 512, 'resnet152': 512}.get(arch, 256)


def do_run(args, classifier=None):
    import torch
    from . import ifcb_io, results, sharding
    from .engine import BinClassifier
    from .neuston_models import NeustonModel
    from .preprocess import parse_imgnorm

    if args.filter:
        if args.filter[0] not in ['IN', 'OUT']:
            raise argparse.ArgumentTypeError('IN|OUT must be either "IN" or "OUT"')
        if len(args.filter) < 2:
            raise argparse.ArgumentTypeError('Must be at least one KEYWORD')
    if classifier is None:
        classifier = NeustonModel.load_from_checkpoint(args.MODEL)
    hp = classifier.hparams
    torch.manual_seed(getattr(hp, 'seed', 0) or 0)
    if os.path.isdir(args.SRC) and not args.SRC.endswith(os.sep):
        args.SRC = args.SRC + os.sep
    if not args.outfile:
        args.outfile = ['D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5'] if args.src_type == 'bin' else ['img_results.json']

    rank, world, local_rank = sharding.env_rank_world()
    if not torch.cuda.is_available():
        raise RuntimeError('RUN needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    filter_mode, keywords = None, []
    if args.filter:
        filter_mode = args.filter[0]
        for kw in args.filter[1:]:
            if os.path.isfile(kw):
                with open(kw) as f:
                    keywords.extend(f.read().splitlines())
            else:
                keywords.append(kw)
    if args.src_type == 'img':
        return _run_images(args, classifier, filter_mode, keywords, rank, world, local_rank)
    if os.path.isdir(args.SRC):
        root = args.SRC
        dd = ifcb_io.DataDirectory(root, whitelist=keywords if filter_mode == 'IN' else None,
                                   blacklist=keywords if filter_mode == 'OUT' else None)
    elif os.path.isfile(args.SRC) and args.SRC.endswith('.txt'):
        with open(args.SRC) as f:
            bins = f.read().splitlines()
        root = os.path.commonpath(bins)
        dd = ifcb_io.DataDirectory(root, whitelist=bins)
    else:
        root = os.path.dirname(args.SRC)
        dd = ifcb_io.DataDirectory(root, whitelist=[os.path.basename(args.SRC)])

    todo = []
    for base in dd.basepaths():
        pid = ifcb_io.Pid(os.path.basename(base))
        pid.namespace = os.path.dirname(base.replace(args.SRC, '')) + os.sep
        if filter_mode == 'IN' and not any(k in str(pid) for k in keywords):
            continue
        if filter_mode == 'OUT' and any(k in str(pid) for k in keywords):
            continue
        if not args.clobber:
            outs = [results.outfile_path(args.outdir, of, pid) for of in args.outfile]
            if all(os.path.isfile(o) for o in outs):
                if rank == 0:
                    print('{} result-file(s) already exist - skipping this bin'.format(pid))
                continue
        todo.append((base, pid))
    mine = set(sharding.my_bins([b for b, _ in todo], rank, world))

    img_norm = parse_imgnorm(hp.img_norm) if getattr(hp, 'img_norm', None) else None
    eng = BinClassifier(hp.MODEL, classifier.model.state_dict(), img_norm=img_norm,
                        transform_input=classifier.model.transform_input, device=torch.device('cuda', local_rank),
                        batch_cap=_engine_batch(args, hp.MODEL), dtype=args.dtype)
    # Host pipeline around the GPU: bin ingest (.roi read straight into a slot of a PINNED ring + C++ .adc parse) runs
    # `depth` bins ahead on I/O threads, the upload is an asynchronous DMA from the slot, and the result files are written
    # behind the GPU by the same pool with a bounded number of writes in flight (each holds one score matrix).
    import collections
    import concurrent.futures as cf
    import queue
    error_bins, n_bins, n_rois, t0 = [], 0, 0, time.time()
    work = [(base, pid) for base, pid in todo if base in mine]
    depth = 2
    n_writers = max(2, min(8, (os.cpu_count() or 4) // max(world, 1) - args.loaders))     # gzip-ing a bin's .h5 / .mat takes about as long as the GPU needs for it
    max_pending = 3 * n_writers
    orientation = os.environ.get('IFCB_ROI_ORIENTATION', 'hw')
    ring = queue.Queue()
    for _ in range(depth + 2):                     # prefetched bins + the one in flight + the one being filed
        ring.put([torch.empty(32 << 20, dtype=torch.uint8).pin_memory()])

    def load(base, pid):
        slot = ring.get()
        try:
            def into(nbytes):
                if nbytes + 16 > slot[0].numel():
                    slot[0] = torch.empty(int(nbytes * 1.25) + 16, dtype=torch.uint8).pin_memory()
                return slot[0].numpy()
            rb = ifcb_io.RawBin(base, into=into, orientation=orientation)
            rb.pid.namespace = pid.namespace
            rb.roi_t = slot[0][:rb.roi.shape[0]]              # the pinned tensor behind rb.roi
            return rb, slot
        except Exception:
            ring.put(slot)
            raise

    def save(pids, bin_pid, scores, top1):
        for of in args.outfile:
            results.save_run_results(pids, scores, hp.classes, args.cmd_timestamp, args.outdir, of,
                                     getattr(hp, 'model_id', None), bin_pid, output_classes=top1)

    def collect(item):
        nonlocal n_bins, n_rois
        name, n, w = item
        try:
            w.result()
            n_bins += 1
            n_rois += n
        except Exception as e:
            error_bins.append((name, type(e).__name__, str(e)))

    pool = cf.ThreadPoolExecutor(max_workers=max(2, args.loaders))          # ingest
    wpool = cf.ThreadPoolExecutor(max_workers=n_writers)                     # result files
    ahead, writes, nxt = collections.deque(), collections.deque(), 0
    inflight = None                                # (ticket, RawBin, ring slot) of the bin the GPU is working on

    def finish(item):
        """Waits for a submitted bin and hands its results to the writer pool; frees its ring slot."""
        ticket, rb, slot = item
        try:
            scores, top1 = eng.fetch(ticket)
            writes.append((str(rb.pid), len(rb), wpool.submit(save, rb.pids, rb.pid, scores.copy(), top1.copy())))
        except Exception as e:
            error_bins.append((str(rb.pid), type(e).__name__, str(e)))
        finally:
            ring.put(slot)                         # the event has fired: the DMA out of the slot is complete
        while len(writes) > max_pending:
            collect(writes.popleft())

    for base, pid in work:
        while len(ahead) < depth and nxt < len(work):
            ahead.append(pool.submit(load, *work[nxt]))
            nxt += 1
        slot = None
        try:
            rb, slot = ahead.popleft().result()
            if len(rb) == 0:
                error_bins.append((str(pid), 'AssertionError', 'Bin is Empty'))
                ring.put(slot)
                continue
            ticket = eng.submit(rb.roi_t, rb.offsets, rb.heights, rb.widths)     # queued behind the bin in flight
        except Exception as e:      # per-bin isolation, as the reference
            error_bins.append((str(pid), type(e).__name__, str(e)))
            if slot is not None:
                ring.put(slot)
            continue
        if inflight is not None:
            finish(inflight)        # the GPU already holds the next bin's work while the host files this one
        inflight = (ticket, rb, slot)
    if inflight is not None:
        finish(inflight)
    while writes:
        collect(writes.popleft())
    pool.shutdown()
    wpool.shutdown()
    summary = dict(rank=rank, n_bins=n_bins, n_rois=n_rois, seconds=time.time() - t0, error_bins=error_bins)
    allsum = sharding.gather_summary(summary, world)
    if rank == 0:
        print('RUN IS DONE')
        tot_b, tot_r = sum(s['n_bins'] for s in allsum), sum(s['n_rois'] for s in allsum)
        print('%d bins, %d ROIs on %d GPU(s) in %.1f s' % (tot_b, tot_r, world, max(s['seconds'] for s in allsum)))
        errs = [e for s in allsum for e in s['error_bins']]
        if errs:
            print('The following bins failed; they were not processed:')
            for b, t, m in errs:
                print(b, t, m)
    return allsum


def _run_images(args, classifier, filter_mode, keywords, rank, world, local_rank):
    """RUN --type img (reference neuston_net.py:280-308): classify image files (a directory tree, a .txt list or one
    file) through the same fused preprocess + network path as bins; images are decoded to gray planes on host
    threads and packed into one byte buffer per batch."""
    import concurrent.futures as cf
    import numpy as np
    import torch
    from . import results
    from .engine import BinClassifier
    from .neuston_data import IMG_EXTENSIONS, load_gray
    hp = classifier.hparams
    ok_ext = lambda p: p.endswith(IMG_EXTENSIONS)             # case-sensitive, as the reference (neuston_net.py:283-292)
    paths = []
    if os.path.isdir(args.SRC):
        for pardir, _, imgs in os.walk(args.SRC):
            paths.extend(os.path.join(pardir, i) for i in sorted(imgs) if ok_ext(i))
    elif os.path.isfile(args.SRC) and args.SRC.endswith('.txt'):
        with open(args.SRC) as f:
            paths = [l.strip() for l in f.read().splitlines() if ok_ext(l.strip())]
    elif ok_ext(args.SRC):
        paths.append(args.SRC)
    if filter_mode == 'IN':
        paths = [p for p in paths if any(k in p for k in keywords)]
    elif filter_mode == 'OUT':
        paths = [p for p in paths if not any(k in p for k in keywords)]
    assert len(paths) > 0, 'No images to process'
    all_paths = paths
    paths = paths[rank::world] if world > 1 else paths          # multi-GPU: each rank takes a stride of the list
    # The reference's ImageDataset (neuston_data.py:386-388) is Resize + ToTensor only -- NO Normalize, even for a model
    # trained with --img-norm; kept as is so that scores match upstream (recorded in DESIGN.md).
    B = _engine_batch(args, hp.MODEL)
    eng = BinClassifier(hp.MODEL, classifier.model.state_dict(), img_norm=None, transform_input=classifier.model.transform_input,
                        device=torch.device('cuda', local_rank), batch_cap=B, dtype=args.dtype, max_rois=B)
    scores = []
    with cf.ThreadPoolExecutor(max_workers=max(1, args.loaders)) as pool:
        for i in range(0, len(paths), B):
            imgs = list(pool.map(load_gray, paths[i:i + B]))
            hs = np.array([im.shape[0] for im in imgs], np.int32)
            ws = np.array([im.shape[1] for im in imgs], np.int32)
            sizes = hs.astype(np.int64) * ws
            offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
            roi = np.concatenate([im.reshape(-1) for im in imgs])
            s, _ = eng.classify_bin(roi, offs, hs, ws)
            scores.append(s.copy())
    scores = np.concatenate(scores) if scores else np.zeros((0, eng.n_classes), np.float32)
    if world > 1:                                                  # one result file, as the reference: rank 0 re-interleaves the strides
        parts = [None] * world
        torch.distributed.all_gather_object(parts, scores)
        if rank != 0:
            return []
        scores = np.zeros((len(all_paths), eng.n_classes), np.float32)
        for r, part in enumerate(parts):
            scores[r::world] = part
        paths = all_paths
    outs = []
    for of in args.outfile:
        outs.append(results.save_run_results(paths, scores, hp.classes, args.cmd_timestamp, args.outdir, of, getattr(hp, 'model_id', None),
                                             args.SRC))
    print('RUN IS DONE')
    return outs


def do_training(args):
    """TRAIN (reference neuston_net.py:37-160): see train_loop.do_training."""
    from .train_loop import do_training as _do_training
    return _do_training(args)


def main(argv=None):
    parser = argparse_nn()
    args = parser.parse_args(argv)
    if args.cmd_mode == 'RUN':
        from .neuston_models import NeustonModel
        classifier = NeustonModel.load_from_checkpoint(args.MODEL)
        argparse_nn_runtimeparams(args, classifier)
        do_run(args, classifier)
    elif args.cmd_mode == 'TRAIN':
        argparse_nn_runtimeparams(args)
        do_training(args)
    else:
        parser.print_help()
        return 2
    return 0


if __name__ == '__main__':
    sys.exit(main())
