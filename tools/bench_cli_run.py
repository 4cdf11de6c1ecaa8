"""File-to-file RUN throughput: `neuston_net RUN` over synthetic bins ON DISK (wall clock; .adc parse, .roi read into the pinned
ring, upload, GPU path, result files -- everything a user's run pays for), one process per GPU with the bins sharded by rank.

    python tools/bench_cli_run.py --bins-per-gpu 64                                    # 1 GPU, default .h5 outputs
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cli_run.py --bins-per-gpu 64

Prints one line per output format: aggregate ROI/s = all ROIs / slowest rank's classify loop (engine construction excluded,
reported separately), and the host's CPU utilisation over the loop (psutil, all cores).
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument('--bins-per-gpu', type=int, default=64)
ap.add_argument('--rois', type=int, default=2048)
ap.add_argument('--model', default='inception_v3')
ap.add_argument('--outfile', action='append', help="default: the reference's default D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5, then .mat, then .json")
ap.add_argument('--loaders', type=int, default=4)
ap.add_argument('--dir', default=None, help='where the synthetic bins are written (default: a temp dir; /dev/shm keeps them off the disk)')
a = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def _write(job):
    from oracle import synth_bins
    idx, rois, root = job
    synth_bins.write_bin(root, synth_bins.make_bin(idx, n_rois=rois))
    return idx


# every rank writes its share of the bins with a few worker processes, BEFORE any CUDA / NCCL state exists (fork)
root = a.dir or os.path.join(tempfile.gettempdir(), 'ifcb_cli_bench_%d' % os.getppid())
bins_dir = os.path.join(root, 'bins')
os.makedirs(bins_dir, exist_ok=True)
total_bins = a.bins_per_gpu * world
mine = [(i, a.rois, bins_dir) for i in range(total_bins) if i % world == rank]
import multiprocessing as mp
t_gen = time.perf_counter()
with mp.get_context('fork').Pool(max(1, min(8, (os.cpu_count() or 1) // world))) as pool:
    pool.map(_write, mine)
t_gen = time.perf_counter() - t_gen

import psutil
import torch
import torch.distributed as dist
from ifcb_classifier_b200 import neuston_net
from ifcb_classifier_b200.neuston_models import NeustonModel

torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('NCCL_DEBUG', 'WARN')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dist.barrier()
ckpt = os.path.join(root, 'bench.ptl')
if rank == 0:
    torch.manual_seed(0)
    hp = argparse.Namespace(MODEL=a.model, classes=['class_%03d' % i for i in range(100)], pretrained=False,
                            resize=299 if a.model == 'inception_v3' else 224, img_norm=None, model_id='bench', seed=1)
    NeustonModel(hp).save_checkpoint(ckpt)
if world > 1:
    dist.barrier()
for outfile in (a.outfile or ['D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5', 'mat/{BIN_ID}_class.mat', 'json/{BIN_ID}_class.json']):
    argv = ['--loaders', str(a.loaders), 'RUN', bins_dir, ckpt, 'R', '--outdir', os.path.join(root, 'out'), '--outfile', outfile, '--clobber']
    args = neuston_net.argparse_nn().parse_args(argv)
    clf = NeustonModel.load_from_checkpoint(args.MODEL)
    neuston_net.argparse_nn_runtimeparams(args, clf)
    if world > 1:
        dist.barrier()
    psutil.cpu_percent(interval=None)
    t0 = time.perf_counter()
    summ = neuston_net.do_run(args, clf)                    # all ranks' summaries (gathered at the end of the run)
    wall = time.perf_counter() - t0
    cpu = psutil.cpu_percent(interval=None)
    if rank == 0:
        n = sum(s['n_rois'] for s in summ)
        loop = max(s['seconds'] for s in summ)
        errs = sum(len(s['error_bins']) for s in summ)
        print(json.dumps(dict(tool='bench_cli_run', model=a.model, outfile=outfile, n_gpus=world, bins=sum(s['n_bins'] for s in summ), rois=n,
                              loop_seconds=loop, rois_per_s=n / loop, wall_seconds_with_engine_build=wall, errors=errs,
                              host_cpu_percent=cpu, host_cores=os.cpu_count(), loaders_per_rank=a.loaders, bin_write_seconds=t_gen,
                              per_rank_rois_per_s=[s['n_rois'] / max(s['seconds'], 1e-9) for s in summ])), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    import shutil
    shutil.rmtree(root, ignore_errors=True)
