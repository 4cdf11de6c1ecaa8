"""``neuston_net.py TRAIN`` re-hosted on the B200 train step (reference neuston_net.py:37-160).

What pytorch_lightning's ``Trainer.fit`` did around ``NeustonModel`` is restated here as a plain loop:
seeded split, shuffled per-rank batches (DistributedSampler semantics), ``TrainNet.step`` per batch,
eval-mode validation through the RUN plan, ``val_loss`` = SUM of the per-batch cross-entropy means
(neuston_models.py:110), best-model checkpoint on ``val_loss``, EarlyStopping(patience=--estop) after
--emin epochs, ``epochs.csv`` / ``args.yml`` / ``training_images.list`` / ``validation_images.list`` /
``{model_id}.ptl`` / validation results files (neuston_callbacks.py:20-156).

Images are decoded on host threads (PIL) to uint8 gray planes, packed into ONE pinned byte buffer per
batch and transformed on the GPU by the fused preprocess kernel -- the same kernel as RUN, so training
and inference see bit-identical inputs.
"""
import concurrent.futures as cf
import os
import random
import time

import numpy as np
import torch

from . import preprocess as pp
from .neuston_data import get_trainval_datasets, load_gray


class ImageBatcher(object):
    """Batches of a NeustonDataset as device tensors: x float32 [B,3,R,R] (written into ``out``), labels, paths."""

    def __init__(self, dataset, batch, device, loaders=4, rank=0, world=1, shuffle=False, seed=0, drop_last=False):
        self.ds, self.batch, self.device = dataset, int(batch), torch.device(device)
        self.rank, self.world, self.shuffle, self.seed, self.drop_last = rank, world, shuffle, seed, drop_last
        self.tf = dataset.transforms or dict(resize=224, img_norm=None, flips=[])
        self.pool = cf.ThreadPoolExecutor(max_workers=max(1, loaders))
        self.epoch = 0

    def indices(self):
        n = len(self.ds)
        idx = list(range(n))
        if self.shuffle:
            random.Random(self.seed * 100003 + self.epoch).shuffle(idx)
        if self.world > 1:                                       # DistributedSampler: pad to a multiple of world, stride by rank
            total = ((n + self.world - 1) // self.world) * self.world
            idx = (idx + idx[:total - n])[self.rank:total:self.world]
        return idx

    def __len__(self):
        n = len(self.indices())
        return n // self.batch if self.drop_last else (n + self.batch - 1) // self.batch

    def _load(self, i, rng_seed):
        img = load_gray(self.ds.images[i])
        r = random.Random(rng_seed)
        for f in self.tf['flips']:                               # RandomVerticalFlip / RandomHorizontalFlip, p = 0.5
            if r.random() < 0.5:
                img = img[::-1] if f == 'v' else img[:, ::-1]
        return np.ascontiguousarray(img)

    def __iter__(self):
        idx = self.indices()
        R = self.tf['resize']
        for b0 in range(0, len(idx), self.batch):
            chunk = idx[b0:b0 + self.batch]
            if self.drop_last and len(chunk) < self.batch:
                break
            imgs = list(self.pool.map(lambda i: self._load(i, (self.seed, self.epoch, i).__hash__()), chunk))
            hs = np.array([im.shape[0] for im in imgs], np.int32)
            ws = np.array([im.shape[1] for im in imgs], np.int32)
            sizes = hs.astype(np.int64) * ws
            offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
            packed = torch.empty(int(sizes.sum()) + 16, dtype=torch.uint8).pin_memory()
            pk = packed.numpy()
            for im, o, s in zip(imgs, offs, sizes):
                pk[o:o + s] = im.reshape(-1)
            d = self.device
            x = pp.preprocess_rois(packed.to(d, non_blocking=True), torch.from_numpy(offs).to(d), torch.from_numpy(hs).to(d),
                                   torch.from_numpy(ws).to(d), R, img_norm=self.tf['img_norm'], out_mode=pp.OUT_F32_NCHW,
                                   max_h=int(hs.max()), max_w=int(ws.max()))
            y = torch.tensor([self.ds.targets[i] for i in chunk], dtype=torch.int64, device=d)
            yield x, y, [self.ds.images[i] for i in chunk]
        self.epoch += 1


def _validation_stats(input_classes, output_classes, n_classes):
    from sklearn import metrics
    idxs = list(range(n_classes))
    stats = {}
    for mode in ['weighted', 'macro', None]:
        for stat in ['f1', 'recall', 'precision']:
            stats['{}_{}'.format(stat, mode if mode else 'perclass')] = getattr(metrics, stat + '_score')(
                input_classes, output_classes, labels=idxs, average=mode, zero_division=0)
    stats['confusion_matrix'] = metrics.confusion_matrix(input_classes, output_classes, labels=idxs, normalize=None)
    return stats


def save_validation_results(outfile, series, args, epoch, train_ds, val_ds, input_classes, outputs, input_srcs):
    """SaveValidationResults.on_validation_end (neuston_callbacks.py:20-156): .mat / .json / .h5 outputs."""
    import json
    labels = args.classes
    input_classes = np.asarray(input_classes)
    output_classes = np.argmax(outputs, axis=1)
    stats = _validation_stats(input_classes, output_classes, len(labels))
    base = lambda p: os.path.splitext(os.path.basename(p))[0]
    vc, tc = val_ds.count_perclass, train_ds.count_perclass
    optional = dict(image_fullpaths=list(input_srcs), image_basenames=[base(p) for p in input_srcs],
                    training_image_fullpaths=list(train_ds.images), training_image_basenames=[base(p) for p in train_ds.images],
                    training_classes=list(train_ds.targets), output_winscores=np.max(outputs, axis=1), output_scores=outputs,
                    counts_perclass=[a + b for a, b in zip(vc, tc)], val_counts_perclass=vc, train_counts_perclass=tc, **stats)
    for stat in ['f1', 'recall', 'precision']:
        optional['classes_by_' + stat] = sorted(range(len(labels)), key=lambda i: stats[stat + '_perclass'][i], reverse=True)
    optional['classes_by_count'] = sorted(range(len(labels)), key=lambda i: optional['counts_perclass'][i], reverse=True)
    res = dict(model_id=args.model_id, timestamp=args.cmd_timestamp, class_labels=labels, input_classes=input_classes,
               output_classes=output_classes)
    res.update({k: v for k, v in optional.items() if k in series and k != 'train_counts_perclass'})
    if 'train_counts_perclass' in series:            # upstream writes the VALIDATION counts under this request (neuston_callbacks.py:100)
        res['val_counts_perclass'] = vc
    outfile = os.path.join(args.outdir, outfile).format(epoch=epoch)
    os.makedirs(os.path.dirname(outfile) or '.', exist_ok=True)
    if outfile.endswith('.json'):
        with open(outfile, 'w') as f:
            json.dump({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in res.items()}, f)
    elif outfile.endswith('.mat'):
        from scipy.io import savemat
        idx_data = ['input_classes', 'output_classes', 'training_classes'] + ['classes_by_' + s for s in 'f1 recall precision count'.split()]
        str_data = ['class_labels', 'image_fullpaths', 'image_basenames', 'training_image_fullpaths', 'training_image_basenames']
        out = {}
        for k, v in res.items():                                   # same dispatch ORDER as upstream (:129-134): arrays first, so the
            if isinstance(v, np.ndarray):                          # 0-based input/output_classes arrays are stored as float32 and only
                out[k] = v.astype('f4')                            # list-typed index series get the 1-based MATLAB shift
            elif isinstance(v, np.float64):
                out[k] = v.astype('f4')
            elif k in str_data:
                out[k] = np.asarray(v, dtype='object')
            elif k in idx_data:
                out[k] = np.asarray(v).astype('u4') + 1
            else:
                out[k] = v
        savemat(outfile, out, do_compression=True)
    elif outfile.endswith('.h5'):
        # _save_validation_results_hdf (neuston_callbacks.py:139-156) through the in-tree HDF5 writer
        from . import h5lite
        attrib = ['model_id', 'timestamp'] + 'f1_weighted recall_weighted precision_weighted f1_macro recall_macro precision_macro'.split()
        int_data = ['input_classes', 'output_classes', 'training_classes', 'counts_perclass', 'val_counts_perclass', 'train_counts_perclass'] + \
                   ['classes_by_' + s for s in 'f1 recall precision count'.split()]
        str_data = ['class_labels', 'image_fullpaths', 'image_basenames', 'training_image_fullpaths', 'training_image_basenames']
        attrs, ds = {}, {}
        for k, v in res.items():
            if k in attrib:
                attrs[k] = (v if v is not None else '') if isinstance(v, (str, type(None))) else float(v)
            elif k in str_data:
                ds[k] = dict(data=[str(s) for s in v], dtype='vlen_str')
            elif k in int_data:
                ds[k] = dict(data=np.asarray(v), dtype='int16')
            elif isinstance(v, np.ndarray):
                ds[k] = dict(data=v, dtype='float16')
            else:
                raise UserWarning('hdf results: WE MISSED THIS ONE: {}'.format(k))
        ds['metadata'] = dict(data=h5lite.Empty('f'), attrs=attrs)
        h5lite.write(outfile, ds)
    return outfile


def do_training(args):
    import yaml
    from . import sharding
    from .graph import CompiledNet
    from .neuston_models import NeustonModel
    from .train import TrainNet

    rank, world, local_rank = sharding.env_rank_world()
    if not torch.cuda.is_available():
        raise RuntimeError('TRAIN needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=dev)
    date_str = args.cmd_timestamp.split('T')[0]
    args.model_id = args.model_id.format(TRAIN_DATE=date_str, TRAIN_ID=args.TRAIN_ID)
    os.makedirs(args.outdir, exist_ok=True)
    if not args.result_files:
        args.result_files = ['results.mat training_image_basenames training_classes image_basenames input_classes output_scores '
                             'confusion_matrix counts_perclass f1_perclass f1_weighted f1_macro'.split()]
    if not args.seed:                                             # 0 = pick one and record it (seed_everything(None))
        args.seed = random.SystemRandom().randint(1, 2 ** 31 - 1) if world == 1 else 1
    random.seed(args.seed)
    np.random.seed(args.seed % (2 ** 32))
    torch.manual_seed(args.seed)

    train_ds, val_ds = get_trainval_datasets(args)
    assert train_ds.classes == val_ds.classes
    args.classes = train_ds.classes
    if rank == 0:
        with open(os.path.join(args.outdir, 'training_images.list'), 'w') as f:
            f.write('\n'.join(sorted(train_ds.images)))
        with open(os.path.join(args.outdir, 'validation_images.list'), 'w') as f:
            f.write('\n'.join(sorted(val_ds.images)))

    classifier = NeustonModel(args)                               # random init or torchvision weights, head swapped
    if getattr(args, 'pretrained', False) and rank == 0:
        print('pretrained=True: torchvision weights are taken from the local cache (no network on this host)')
    B = args.batch_size
    net = TrainNet(args.MODEL, classifier.model.state_dict(), B, device=dev, dtype=getattr(args, 'train_dtype', 'bf16'), seed=args.seed,
                   R=args.resize, transform_input=classifier.model.transform_input, deterministic=getattr(args, 'deterministic', False))
    net.enable_cuda_graph()                                       # replay the step's ~10^3 launches from CUDA graphs
    train_loader = ImageBatcher(train_ds, B, dev, args.loaders, rank, world, shuffle=True, seed=args.seed, drop_last=False)
    val_loader = ImageBatcher(val_ds, B, dev, args.loaders, 0, 1, shuffle=False, seed=args.seed)

    chk_dir = os.path.join(args.outdir, 'chkpts')
    os.makedirs(chk_dir, exist_ok=True)
    log_rows, best_val, best_epoch, best_path, wait = [], np.inf, 0, None, 0
    hp = {k: v for k, v in vars(args).items()}
    tail_nets = {}                                                # plans for the short last batch of an epoch (shared arenas)
    ev = None                                                     # eval-mode plan of the validation pass
    for epoch in range(args.emax):
        t0 = time.time()
        agg_train_loss, losses = 0.0, []
        for x, y, _ in train_loader:
            n = int(x.shape[0])
            if n < B:
                # the reference trains on the short last batch as it is (BN statistics and the logged loss are those of the
                # n samples): a second plan of that size over the SAME parameter / Adam arenas
                if n not in tail_nets:
                    tail_nets[n] = TrainNet(args.MODEL, classifier.model.state_dict(), n, device=dev, dtype=getattr(args, 'train_dtype', 'bf16'),
                                            seed=args.seed, R=args.resize, transform_input=classifier.model.transform_input, share=net,
                                            deterministic=getattr(args, 'deterministic', False))
                tail = tail_nets[n]
                tail.repack()                                     # its 16-bit operands are stale: the main plan has stepped since
                losses.append(tail.step(x, y).clone())
                net.repack()
                continue
            losses.append(net.step(x, y).clone())
        agg_train_loss = float(torch.stack(losses).sum()) if losses else 0.0       # one sync per epoch (reference: .item() per step)
        # ---- validation (eval mode = the RUN plan with the current weights).  Data parallel: every rank holds the same
        # parameters but its own BatchNorm running statistics; rank 0's are broadcast first (DDP broadcast_buffers), so every
        # rank validates the same model, and rank 0's val_loss decides best / early stop for all ranks (no rank can leave the
        # epoch loop alone and strand the others in the gradient all-reduce).
        if world > 1:
            for k_, b_ in net.buffers.items():
                if b_.is_cuda:
                    torch.distributed.broadcast(b_, 0)
        sd = net.state_dict()
        if ev is None:                            # one eval plan for the whole run; later epochs refresh its weights in place
            ev = CompiledNet(args.MODEL, sd, B, in_kind='f32', R=args.resize, device=dev, dtype=getattr(args, 'dtype', 'fp16'),
                             transform_input=classifier.model.transform_input)
        else:
            ev.load_state_dict(sd)
        val_loss, outs, ins, srcs = 0.0, [], [], []
        for x, y, paths in val_loader:
            n = int(x.shape[0])
            ev.inp[:n].copy_(x)
            scores, logits, _, _ = ev.forward(n)
            val_loss += float(torch.nn.functional.cross_entropy(logits, y))        # validation_step: CE of the eval-mode logits
            outs.append(scores.cpu().numpy().copy())
            ins.append(y.cpu().numpy())
            srcs.extend(paths)
        outputs, input_classes = np.concatenate(outs), np.concatenate(ins)
        stats = _validation_stats(input_classes, np.argmax(outputs, 1), len(args.classes))
        if world > 1:
            vl = torch.tensor([val_loss], dtype=torch.float64, device=dev)
            torch.distributed.broadcast(vl, 0)
            val_loss = float(vl.item())
        is_best = val_loss < best_val
        if is_best:
            best_val, best_epoch, wait = val_loss, epoch, 0
        else:
            wait += 1
        if rank == 0:
            print('Best Epoch: {}, train_loss: {:.3f}, val_loss: {:.3f}, val_f1_w={:02.1f}%, val_f1_m={:02.1f}% ({:.1f}s)'.format(
                True if is_best else best_epoch + 1, agg_train_loss, val_loss, 100 * stats['f1_weighted'], 100 * stats['f1_macro'],
                time.time() - t0), flush=True)
            log_rows.append(dict(epoch=epoch, best=bool(is_best), train_loss=agg_train_loss, val_loss=val_loss,
                                 f1_macro=float(stats['f1_macro']), f1_weighted=float(stats['f1_weighted'])))
            if is_best:
                classifier.model.load_state_dict(sd)
                classifier.hparams.epoch = epoch
                if best_path and os.path.exists(best_path):
                    os.remove(best_path)
                best_path = os.path.join(chk_dir, 'epoch={}.ckpt'.format(epoch))
                classifier.save_checkpoint(best_path)
                for rf in args.result_files:
                    save_validation_results(rf[0], rf[1:], args, epoch, train_ds, val_ds, input_classes, outputs, srcs)
        if args.estop and wait >= args.estop and epoch + 1 >= args.emin:
            break
    if rank == 0:
        import shutil
        if best_path is None:
            raise RuntimeError('TRAIN: no epoch produced a finite val_loss (last: %r) -- no checkpoint to keep as %s.ptl' % (val_loss, args.model_id))
        shutil.copyfile(best_path, os.path.join(args.outdir, args.model_id + '.ptl'))
        if args.epochs_log:
            cols = ['epoch', 'best', 'train_loss', 'val_loss', 'f1_macro', 'f1_weighted']
            with open(os.path.join(args.outdir, args.epochs_log), 'w') as f:
                f.write(','.join(cols) + '\n')
                for r in log_rows:
                    f.write(','.join(str(r[c]) for c in cols) + '\n')
        if args.args_log:
            with open(os.path.join(args.outdir, args.args_log), 'w') as f:
                yaml.safe_dump({k: (v if isinstance(v, (int, float, str, bool, list, type(None))) else str(v)) for k, v in hp.items()}, f)
        if getattr(args, 'onnx', False):
            print('--onnx: ONNX export is outside the accelerated path (DESIGN.md, out of scope); skipped')
    return dict(best_epoch=best_epoch, best_val_loss=best_val, epochs=len(log_rows) if rank == 0 else None, log=log_rows)
