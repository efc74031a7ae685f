"""Is K1 alone slower when several GPUs of the box run it at the same time?  Launch under torchrun; every rank times
K1 (graph replay, no exchange) on its own pages, first without and then with an initialised NCCL process group.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/k1_multi_gpu_probe.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ["RN_B200_PEER_BOX"] = "0"
import retinanet_b200 as rn  # noqa: E402
import synthetic  # noqa: E402

HW, B = (800, 1333), 16
anchors = rn.anchors_for_shape(HW + (3,))
images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors), first_page=0)    # the SAME pages on every rank
cls, reg = synthetic.training_predictions(2, B, anchors.shape[0], classes=1)


def measure(tag, images=images, anns=anns):
    step = rn.pipeline.TargetLossStep(HW + (3,), B, 22, 1, peer_box=False)
    step.load_annotations(images, anns)
    step.load_predictions(torch.from_numpy(cls), torch.from_numpy(reg))
    step._build_graphs()
    out = []
    for which in (0, 1):
        g = step._graphs[which]
        for _ in range(10):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
            torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 200 * 1e3)
    print("rank %d %-10s K1 %.2f us  K2 %.2f us" % (rank, tag, out[0], out[1]), flush=True)


measure("no-nccl")
# the pages the benchmark gives this rank (dealt out by estimated cost, heaviest first), in that order and in page order
g_images, g_anns = synthetic.training_batch(2, batch=max(world, 2) * B, anchors=np.asarray(anchors), first_page=0)
cost = rn.distributed.page_cost(g_anns, HW)
mine = rn.distributed.balanced_shards(cost, max(world, 2))[rank % max(world, 2)]
print("rank %d shard cost %.1f  pages %s" % (rank, float(np.sum(np.asarray(cost)[mine])), list(mine)), flush=True)
measure("shard", [g_images[i] for i in mine], [g_anns[i] for i in mine])
srt = sorted(mine)
measure("shard-sorted", [g_images[i] for i in srt], [g_anns[i] for i in srt])
time.sleep(1.0)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = torch.ones(1, device="cuda")
    dist.all_reduce(t)
    measure("nccl-up")
    dist.barrier()
    dist.destroy_process_group()
