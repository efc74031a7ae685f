"""CPU, world_size 2, gloo: the N>1 host logic.  Pages shard by image with no data-path collective; the
only exchange is one all-reduce of the positive-anchor count (before gradients are scaled) and one of the
loss sums.  The per-rank arithmetic is done by the oracle here (no GPU), the plumbing under test is
retinanet_b200.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import retinanet_b200 as rn
    import synthetic
    from oracle import anchors_np as O
    from oracle import losses_np as OL
    hw = (128, 160)
    anc = O.anchors_for_shape(hw + (3,))
    pages = 5
    imgs = [synthetic.PageShape(hw + (3,)) for _ in range(pages)]
    anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=4) for i in range(pages)]
    rs = np.random.RandomState(0)
    cls = (1 / (1 + np.exp(-rs.normal(-3, 1, (pages, anc.shape[0], 1))))).astype(np.float32)
    reg = rs.normal(0, 1, (pages, anc.shape[0], 4)).astype(np.float32)
    assert rn.distributed.PeerCounter.create() is None      # no GPU here: the all_reduce fallback is the path under test
    lo, hi = rn.distributed.shard_pages(pages)
    y_reg, y_cls = O.anchor_targets_bbox(anc, imgs[lo:hi], anns[lo:hi], 1)
    npos_local = torch.tensor((y_cls[:, :, -1] == 1).sum(axis=1), dtype=torch.int32)
    npos = rn.distributed.global_positive_count(npos_local)
    lf, gf = OL.focal()(y_cls, cls[lo:hi], return_grad=True, normalizer=max(1.0, float(npos)))
    ls, gs = OL.smooth_l1()(y_reg, reg[lo:hi], return_grad=True, normalizer=max(1.0, float(npos)))
    total = rn.distributed.reduce_losses(torch.tensor([lf, ls], dtype=torch.float32))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, npos=npos.numpy(), total=total.numpy(), gf=gf, gs=gs)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_losses_match_single_process(tmp_path):
    import synthetic
    from oracle import anchors_np as O
    from oracle import losses_np as OL
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    hw = (128, 160)
    anc = O.anchors_for_shape(hw + (3,))
    pages = 5
    imgs = [synthetic.PageShape(hw + (3,)) for _ in range(pages)]
    anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=4) for i in range(pages)]
    rs = np.random.RandomState(0)
    cls = (1 / (1 + np.exp(-rs.normal(-3, 1, (pages, anc.shape[0], 1))))).astype(np.float32)
    reg = rs.normal(0, 1, (pages, anc.shape[0], 4)).astype(np.float32)
    y_reg, y_cls = O.anchor_targets_bbox(anc, imgs, anns, 1)
    lf, gf = OL.focal()(y_cls, cls, return_grad=True)
    ls, gs = OL.smooth_l1()(y_reg, reg, return_grad=True)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(r['lo']) for r in ranks] == [0, 3] and [int(r['hi']) for r in ranks] == [3, 5]
    n_glob = float((y_cls[:, :, -1] == 1).sum())
    for r in ranks:
        assert float(r['npos'][0]) == n_glob                       # same global normaliser on every rank
        assert np.allclose(r['total'], [lf, ls], rtol=1e-5)        # fp32 sum order differs across world sizes
    assert np.allclose(np.concatenate([r['gf'] for r in ranks]), gf, rtol=1e-6, atol=1e-12)
    assert np.allclose(np.concatenate([r['gs'] for r in ranks]), gs, rtol=1e-6, atol=1e-12)
