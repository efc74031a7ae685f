"""Benchmark of the RetinaNet anchor + detection-head path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the training-target path on a batch of 16 pages of 800x1333, 1 class,
<= 20 GT tables per page -- K1 (anchor generation + IoU matching + targets) then K2 (focal + smooth-L1
forward and backward).  One "step" = one batch per GPU; per-GPU work is fixed as N grows (weak scaling),
pages shard by image, the only exchange is the positive-anchor count (NVLink peer mailbox, or NCCL all-reduce).

`value`  = pages/s with the batch's inputs resident in HBM (GT block + head outputs), CUDA events.  Every step launches
           one K1 and one K2 as two branches of one CUDA graph: K1 on the batch loaded now, K2 on the batch before it
           (double-buffered targets; bit-identical to `in_order`, which is K1(s) then K2(s) and reported beside it).
`e2e`    = pages/s through the public Python API (HostStepPipeline.submit / result) with HOST inputs every step: the
           ragged GT list is packed and copied, the head outputs are copied from pinned memory on a copy stream, K1 +
           K2 run, the loss scalars are read back; two steps are in flight so the copy of step s+1 overlaps the
           kernels of step s (`e2e.synchronous` = one step at a time, TargetLossStep.run_from_host).
`roofline` = K1, the dominant kernel of the step by time (instruction-issue bound): algorithmic bytes / its mean
           duration (CUDA events inside the timed region) against the measured HBM peak in MEASURED_PEAKS.json.
`roofline_k2` = the same for K2, the HBM-bound loss kernel (north_star's 60 % target).
`cpu_baseline` = the oracle (numpy port of the reference) on this host's cores, bounded sample.
`inference` (extra) = BASELINE configs[2]: 64 pages, fused decode + clip + threshold + sort + NMS.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import synthetic  # noqa: E402

CFG = 2
HW = synthetic.CONFIGS[CFG]['hw']
PAGES_PER_GPU = synthetic.CONFIGS[CFG]['batch']
GMAX = synthetic.CONFIGS[CFG]['gmax'] + 2          # +2: the adversarial snapped duplicates
CLASSES = 1
E2E_GATHER = True                                  # smooth-L1 reads the positive anchors' regression rows straight from pinned host memory
E2E_DEPTH = 2                                      # host-input steps in flight (HostStepPipeline slots)
E2E_CHUNKS = 1                                     # page chunks of the overlapped host-input step
METRIC = "pages/sec (anchor targets + focal/smooth-L1 fwd+bwd) @800x1333"
WORKLOAD = "configs[1]: training-target path, 16 pages/GPU of 800x1333, 1 class, <=20 GT, 200700 anchors/page"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json, written by hand from profiles/*_ncu_full.md)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(kernel, {}).get("bytes_per_launch")
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during warm-up + timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [n for b, n in self.REASONS.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(inside)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (numpy port of the reference) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_page(args):
    """One page of the training-target path on the CPU: targets + both losses forward and backward."""
    page, = args
    from oracle import anchors_np as OA
    from oracle import losses_np as OL
    anchors = _cpu_page.anchors
    img = synthetic.PageShape(HW + (3,))
    ann = synthetic.gt_for_page(CFG, page, anchors=anchors)
    reg, lab = OA.anchor_targets_bbox(anchors, [img], [ann], CLASSES)
    cls, rp = synthetic.training_predictions(CFG, 1, anchors.shape[0], classes=CLASSES, first_page=page)
    npos = float((lab[:, :, -1] == 1).sum())
    # the batch-global normaliser needs every page's count first; per page the arithmetic is identical,
    # so the worker uses its own count (same work, documented in DESIGN.md)
    lf, gf = OL.focal()(lab, cls, return_grad=True, normalizer=max(1.0, npos))
    ls, gs = OL.smooth_l1()(reg, rp, return_grad=True, normalizer=max(1.0, npos))
    return npos, float(lf), float(ls)


def _cpu_init():
    from oracle import anchors_np as OA
    _cpu_page.anchors = OA.anchors_for_shape(HW + (3,))
    os.environ["OMP_NUM_THREADS"] = "1"


class CpuPool(object):
    def __init__(self):
        import multiprocessing as mp
        self.cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init)

    def run(self, pages):
        t0 = time.perf_counter()
        self.pool.map(_cpu_page, [(p,) for p in pages], chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline_leg(budget_s=20.0):
    pool = CpuPool()
    try:
        pool.run(range(min(pool.cores, PAGES_PER_GPU)))                 # warm-up (imports, anchors)
        pages, elapsed, reps = 0, 0.0, 0
        while elapsed < budget_s * 0.75 and reps < 400:
            elapsed += pool.run(range(PAGES_PER_GPU))
            pages += PAGES_PER_GPU
            reps += 1
        return {"value": pages / elapsed, "unit": "pages/s", "cores": pool.cores, "kind": "port",
                "sample": "%d x the 16-page batch (targets + losses fwd+bwd per page), numpy oracle, "
                          "multiprocessing.Pool(%d), %.1f s" % (reps, pool.cores, elapsed)}
    finally:
        pool.close()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (numpy oracle port: the reference is
    Python/TF and cannot travel to the GPU box) on all host cores; each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool = CpuPool()
    try:
        t_full = pool.run(range(PAGES_PER_GPU))                          # also warms the workers
        budget = 150.0
        per_step = int(max(1, min(PAGES_PER_GPU, PAGES_PER_GPU * budget / max(1e-9, t_full * (args.steps + args.warmup)))))
        for _ in range(args.warmup):
            pool.run(range(per_step))
        t = 0.0
        for _ in range(args.steps):
            t += pool.run(range(per_step))
        value = per_step * args.steps / t
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32",
                "data": "synthetic", "config": {"workload": WORKLOAD, "pages_per_step": per_step},
                "cpu_baseline": {"value": value, "unit": "pages/s", "cores": pool.cores, "kind": "port",
                                 "sample": "%d pages per step through multiprocessing.Pool(%d), numpy oracle port of "
                                           "model/anchors.py + model/losses.py (fwd+bwd)" % (per_step, pool.cores)},
                "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
    finally:
        pool.close()


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    # stdout carries ONE JSON line: everything else written to fd 1 (NCCL prints its version banner there) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    import retinanet_b200 as rn
    rn._lib.load()

    B, C = PAGES_PER_GPU, CLASSES
    anchors = rn.anchors_for_shape(HW + (3,))
    N = anchors.shape[0]
    first = rank * B
    if world > 1 and not args.contiguous_shards:
        # the global batch (world x 16 pages) is sharded by estimated K1 cost so that every rank's K1 takes about the same
        # time (distributed.balanced_shards); every rank derives the same assignment from the annotations
        g_images, g_anns = synthetic.training_batch(CFG, batch=world * B, anchors=np.asarray(anchors), first_page=0)
        mine = rn.distributed.balanced_shards(rn.distributed.page_cost(g_anns, HW), world)[rank]
        images, anns = [g_images[i] for i in mine], [g_anns[i] for i in mine]
        sharding = "pages of the global batch dealt out by estimated K1 cost (distributed.page_cost + balanced_shards)"
    else:
        images, anns = synthetic.training_batch(CFG, batch=B, anchors=np.asarray(anchors), first_page=first)
        sharding = "contiguous page ranges"
    cls_np, reg_np = synthetic.training_predictions(CFG, B, N, classes=C, first_page=first)
    cls_host = torch.from_numpy(cls_np).pin_memory()
    reg_host = torch.from_numpy(reg_np).pin_memory()

    step = rn.pipeline.TargetLossStep(HW + (3,), B, GMAX, C)
    gt_bytes = step.load_annotations(images, anns)
    step.load_predictions(cls_host, reg_host)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident: `value` + per-kernel durations ----------------------------------------------------
    pipelined = world > 1 and step.peer is not None and args.pipelined
    run_step = step.run_pipelined if pipelined else step.run
    for _ in range(max(args.warmup, 3)):
        step.run()
    exchange = "none (1 rank)"
    if world > 1:
        # the normaliser K2 used must be the all-reduced positive count, whichever way it was exchanged
        check = step.npos_total.clone()
        if step.peer is None:
            check = check * 0 + step.losses[2]            # npos_total already holds the all-reduced count
        else:
            dist.all_reduce(check)
        assert float(check) == float(step.losses[2]) or float(check) < 1.0, (float(check), float(step.losses[2]))
        exchange = ("NVLink peer mailbox, " + ("send + wait fused into K2 (CTA 0 stores the count into every rank's mailbox, all CTAs "
                    "wait on local memory)" if step.peer_fused else "rn_peer_publish kernel + K2 prologue")) if step.peer is not None else "NCCL all_reduce"
    # timed region 1 (`value`): K steps of the product call -- step.run() replays ONE graph per step (K1 [+ publish] + K2)
    barrier()
    t_wall0 = time.perf_counter()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    for i in range(args.steps):
        run_step()
    v1.record()
    barrier()
    t_wall1 = time.perf_counter()
    total_ms = v0.elapsed_time(v1)
    # extra: the overlapped schedule (K1 of the next batch concurrently with K2 of this one, two graph branches)
    ov_ms, ov_match = float("nan"), None
    inorder_losses = step.losses.clone()
    if world == 1 or step.peer_fused:
        # (several ranks: K2 of batch s sends and collects the counts of batch s itself while K1 of batch s+1 runs beside
        # it -- fused publish -- so the exchange and the skew between ranks hide behind the longer kernel)
        for _ in range(max(args.warmup, 3)):
            step.run_pipelined(overlap=True)
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for i in range(args.steps):
            step.run_pipelined(overlap=True)
        o1.record()
        barrier()
        ov_ms = o0.elapsed_time(o1)
        ov_match = bool(torch.equal(step.losses, inorder_losses))     # same batch every step -> same bits as in order
        t_wall1 = time.perf_counter()
    # timed region 2 (per-kernel durations for the rooflines): the same K steps with the two halves replayed
    # separately and CUDA events between them (costs one more graph launch per step, so it is not the `value`)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        run_step(events=evs[i])
    barrier()
    split_ms = evs[0][0].elapsed_time(evs[-1][2])
    k1_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    k2_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    losses = step.losses.cpu().numpy()

    # ---- e2e: public API, host inputs every step -------------------------------------------------------
    def e2e_step():
        # public API with HOST inputs: ragged GT (Python dicts) packed + copied, head outputs copied from pinned
        # memory chunk by chunk while K1 runs, K2 per chunk, 4 x 12 bytes of losses read back (synchronises)
        return step.run_from_host(images, anns, cls_host, reg_host, chunks=E2E_CHUNKS, gather_reg_from_host=E2E_GATHER)
    e2e_steps = 0 if args.no_e2e else args.steps
    for _ in range(0 if args.no_e2e else max(args.warmup, 3)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    sync_ms = e0.elapsed_time(e1) if e2e_steps else float("nan")
    # the headline e2e: the same step through HostStepPipeline, two steps in flight (submit step s+1, then take the
    # result of step s) -- every step still copies its inputs from pinned host memory and reads its losses back
    e2e_ms = float("nan")
    if e2e_steps:
        pipe = rn.pipeline.HostStepPipeline(HW + (3,), B, GMAX, C, depth=E2E_DEPTH)

        def pipe_steps(n):
            prev = None
            for _ in range(n):
                k = pipe.submit(images, anns, cls_host, reg_host, chunks=E2E_CHUNKS, gather_reg_from_host=E2E_GATHER)
                if prev is not None:
                    pipe.result(prev)
                prev = k
            return pipe.result(prev)
        pipe_steps(max(args.warmup, 3))
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pipe_losses = pipe_steps(e2e_steps)
        p1.record()
        barrier()
        e2e_ms = p0.elapsed_time(p1)
        sync_losses = step.run_from_host(images, anns, cls_host, reg_host, chunks=E2E_CHUNKS, gather_reg_from_host=E2E_GATHER)
        assert np.array_equal(pipe_losses.numpy(), sync_losses.numpy()), (pipe_losses, sync_losses)
        del pipe
    # bytes that cross PCIe per step: the GT block, the classification tensor, and -- with the gather -- only the
    # 16-byte regression rows of the positive anchors (K2 reads them in place); the full-copy variant is timed too
    n_pos_rank = float(step.npos.sum().item())
    h2d = gt_bytes + cls_host.numel() * 4 + (16 * n_pos_rank if E2E_GATHER else reg_host.numel() * 4)
    full_ms = float("nan")
    if E2E_GATHER and not args.no_e2e:
        full_step = lambda: step.run_from_host(images, anns, cls_host, reg_host, chunks=E2E_CHUNKS)
        for _ in range(3):
            full_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            full_step()
        f1.record()
        barrier()
        full_ms = f0.elapsed_time(f1)
    sampler.stop_flag = True

    times = torch.tensor([total_ms, e2e_ms, k1_ms, k2_ms, split_ms, full_ms, sync_ms, ov_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, k1_ms, k2_ms, split_ms, full_ms, sync_ms, ov_ms = [float(x) for x in times.cpu()]
    # `value`: the overlapped schedule (K1 of batch s+1 beside K2 of batch s: one K1 + one K2 per step, one graph launch)
    # unless --schedule in-order or the schedule is unavailable (several ranks without the peer mailbox)
    overlapped = args.schedule == "overlapped" and ov_ms == ov_ms and not pipelined
    value_ms = ov_ms if overlapped else total_ms

    # ---- N2 (extra object): K2 fed by the per-level head outputs, sigmoid fused ---------------------------
    levels = None
    if not args.no_inference:
        spec = step.spec
        rows = [int(h) * int(w) * spec.per_cell for h, w in spec.level_hw]
        rs = np.random.RandomState(77 + rank)
        cls_l = [torch.from_numpy(rs.normal(-4.6, 1.0, (B, n, 1)).astype(np.float32)).to(device) for n in rows]
        reg_l = [torch.from_numpy(rs.normal(0.0, 1.0, (B, n, 4)).astype(np.float32)).to(device) for n in rows]
        outs = (torch.empty(3, dtype=torch.float32, device=device), [torch.empty_like(c) for c in cls_l], [torch.empty_like(r) for r in reg_l])
        run_levels = lambda: rn.detection_losses_levels(step.y_reg, step.y_cls, reg_l, cls_l, normalizer=step.npos_total,
                                                        from_logits=True, out=outs, workspace=step.loss_ws)
        for _ in range(5):
            run_levels()
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        l0.record()
        for _ in range(reps):
            run_levels()
        l1.record()
        torch.cuda.synchronize()
        lv_us = l0.elapsed_time(l1) / reps * 1e3
        levels = {"kernel": "k_loss_c1_levels (per-level logits in, sigmoid fused, per-level gradients out)",
                  "us_per_launch": lv_us, "level_rows": rows,
                  "note": "back-to-back launches incl. launch overhead; replaces sigmoid + Concatenate(axis=1) + K2"}

    # ---- inference path (extra object) -----------------------------------------------------------------
    inference = None
    if not args.no_inference:
        inference = bench_inference(rn, torch, device, rank, world, args)

    if rank == 0:
        peak, peak_src = measured_peak()
        # K2 algorithmic bytes: per anchor read labels 4(C+1) + probabilities 4C, write
        # gradients 4C + 16; the anchor state comes from the label row (shared_state: both target tensors are
        # K1's) so regression rows (20 B targets + 16 B prediction) are only read for positive anchors.
        # SURVEY 8(d) counts 12C + 56 = 68 B/anchor (every regression row read); given for comparison.
        n_pos = float(losses[2])
        k2_bytes = (12 * C + 20) * N * B + 36 * n_pos
        k2_bytes_survey = (12 * C + 56) * N * B
        k1_bytes = 4 * (5 + C + 1) * N * B            # 28 B/anchor at C=1
        achieved = k2_bytes / (k2_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * B * args.steps / (value_ms * 1e-3), "unit": "pages/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": value_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 matching + f32 targets/losses", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pages_per_gpu": B, "anchors_per_page": N, "classes": C,
                       "l2": "working set per step ~%d MB (> 126 MB L2), no explicit flush" % ((k1_bytes + k2_bytes) // (1 << 20)),
                       "cuda_graphs": True, "count_exchange": exchange, "sharding": sharding,
                       "schedule": ("overlapped: every step launches K1 of batch s+1 and K2 of batch s as two branches of one "
                                    "graph (TargetLossStep.run_pipelined(overlap=True), double-buffered targets; bit-identical "
                                    "to in order)") if overlapped else
                                   ("pipelined: K1 of batch s+1 enqueued ahead of K2 of batch s (double-buffered targets)"
                                    if pipelined else "in order: K1(s), K2(s)")},
            "in_order": {"pages_per_s": world * B * args.steps / (total_ms * 1e-3), "ms_per_step": total_ms / args.steps,
                         "note": "TargetLossStep.run(): K1(s) then K2(s), one graph launch per step"},
            "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": "pages/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 12 * E2E_CHUNKS, "ms_per_step": e2e_ms / args.steps,
                    "h2d_GBps": h2d / (e2e_ms / args.steps * 1e-3) / 1e9,
                    "api": "HostStepPipeline.submit / result (%d steps in flight: the copy of step s+1 runs during K1 / K2 of "
                           "step s; per step: GT packed + copied, classification tensor copied from pinned memory, K1, K2, "
                           "loss rows read back)" % E2E_DEPTH,
                    "synchronous": {"value": world * B * args.steps / (sync_ms * 1e-3), "ms_per_step": sync_ms / args.steps,
                                    "api": "TargetLossStep.run_from_host (one step at a time, host waits for the losses)"},
                    "regression_rows": ("positive anchors' rows read in place from pinned host memory (model/losses.py:72-74 "
                                        "gathers exactly those); the (B,N,4) tensor is not copied") if E2E_GATHER else "copied",
                    "full_copy": {"value": world * B * args.steps / (full_ms * 1e-3), "ms_per_step": full_ms / args.steps,
                                  "h2d_bytes_per_step": int(gt_bytes + cls_host.numel() * 4 + reg_host.numel() * 4)}},
            "gpu_launches": step.kernel_launches_per_step * args.steps,   # `value` region: K1 + K2 per step (+ publish when not fused into K2)
            # the dominant kernel of the step by time is K1 (~74 %, profiles/*_launches_value_region.md): it is
            # reported first although it is instruction-issue bound, not HBM bound; K2 (the HBM-bound loss kernel
            # north_star sets the 60 % target for) follows
            "roofline": {"kernel": "k_anchor_targets_tiles (K1 anchors + IoU/argmax matching + targets)", "bound": "hbm",
                         "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic("k_anchor_targets_tiles"),
                         "peak_source": peak_src, "bytes_per_launch": k1_bytes, "bytes_per_anchor": k1_bytes / (N * B),
                         "us_per_launch": k1_ms * 1e3, "share_of_step": k1_ms / (k1_ms + k2_ms),
                         "note": "write-only 28 B/anchor; limited by instruction issue (fp64 matching for every anchor x "
                                 "overlapping GT; ncu: issue slots 74 % busy, DRAM 6 %), see DESIGN.md section 3"},
            "roofline_k2": {"kernel": "k_loss_c1_fast (K2 fused focal + smooth-L1 fwd+bwd)", "bound": "hbm",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": ncu_traffic("k_loss_c1_fast"), "peak_source": peak_src, "bytes_per_launch": k2_bytes,
                            "bytes_per_anchor": k2_bytes / (N * B), "us_per_launch": k2_ms * 1e3,
                            "share_of_step": k2_ms / (k1_ms + k2_ms),
                            "achieved_survey_bytes": k2_bytes_survey / (k2_ms * 1e-3) / 1e9},
            "ms_per_step_split_graphs": split_ms / args.steps,
            "overlapped_schedule": None if ov_ms != ov_ms else {
                "pages_per_s": world * B * args.steps / (ov_ms * 1e-3), "ms_per_step": ov_ms / args.steps,
                "losses_equal_in_order": ov_match,
                "note": "TargetLossStep.run_pipelined(overlap=True): K1 of batch s+1 and K2 of batch s as two branches of one graph"
                        + ("; K2 sends and collects the ranks' positive counts itself (fused publish), hidden behind K1" if world > 1 else "")},
            "kernels": {"K1_anchor_targets": {"us": k1_ms * 1e3, "algorithmic_bytes": k1_bytes,
                                              "GBps": k1_bytes / (k1_ms * 1e-3) / 1e9},
                        "K2_losses": {"us": k2_ms * 1e3, "algorithmic_bytes": k2_bytes, "GBps": achieved}},
            "losses": {"focal": float(losses[0]), "smooth_l1": float(losses[1]), "normalizer": float(losses[2])},
            "clocks": sampler.summary(t_wall0, t_wall1),
        }
        if inference is not None:
            line["inference"] = inference
        if levels is not None:
            levels["GBps"] = k2_bytes / (levels["us_per_launch"] * 1e-6) / 1e9
            line["per_level_heads"] = levels
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg()
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_inference(rn, torch, device, rank, world, args):
    """BASELINE configs[2]: 64 pages/GPU, score > 0.05, NMS 0.5, 300 detections, fused head."""
    import torch.distributed as dist
    cfg, B = 3, synthetic.CONFIGS[3]['batch']
    anchors = np.asarray(rn.anchors_for_shape(HW + (3,)))
    N = anchors.shape[0]
    _, anns = synthetic.training_batch(cfg, batch=B, first_page=rank * B)
    cls_np, reg_np = synthetic.inference_predictions(cfg, B, anchors, anns, classes=1, first_page=rank * B)
    cls_host, reg_host = torch.from_numpy(cls_np).pin_memory(), torch.from_numpy(reg_np).pin_memory()
    cls_d, reg_d = cls_host.to(device), reg_host.to(device)
    shape = (B,) + HW + (3,)
    out = {}
    for tag, topk in (("reference_semantics", 0), ("pre_nms_top_k_1000", 1000)):
        head = rn.DetectionHead(pre_nms_top_k=topk)
        for _ in range(3):
            head([shape, reg_d, cls_d])
        steps = max(5, min(args.steps, 50))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = float("inf")
        for _ in range(3):                                  # best of 3 timing loops (this leg is an extra, not the headline)
            e0.record()
            for _ in range(steps):
                res = head([shape, reg_d, cls_d])
            e1.record()
            torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1) / steps)
        # two batches in flight on two streams (serving loop): k_segment_nms runs one CTA per (page, class), i.e. 64
        # of the 148 SMs per batch, so a second independent batch fills the other SMs.  Workspaces are per stream.
        streams = [torch.cuda.Stream(device) for _ in range(2)]
        for st in streams:
            st.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(st):
                head([shape, reg_d, cls_d])
        torch.cuda.synchronize()
        e0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream(device))
        for i in range(2 * steps):
            with torch.cuda.stream(streams[i % 2]):
                res2 = head([shape, reg_d, cls_d])
        for st in streams:
            torch.cuda.current_stream(device).wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        ms_2s = e0.elapsed_time(e1) / (2 * steps)
        assert torch.equal(res2[1], res[1])
        # e2e: host head outputs in, detections out
        # (a) both tensors copied; (b) scores copied, the candidates' regression rows read in place from pinned memory
        e0.record()
        for _ in range(steps):
            r = head([shape, reg_host.to(device, non_blocking=True), cls_host.to(device, non_blocking=True)])
            host = [t.cpu() for t in r]
        e1.record()
        torch.cuda.synchronize()
        ms_e2e_copy = e0.elapsed_time(e1) / steps
        r = head([shape, reg_host, cls_host.to(device, non_blocking=True)])
        assert torch.equal(r[0], res[0]) and torch.equal(r[1], res[1])
        e0.record()
        for _ in range(steps):
            r = head([shape, reg_host, cls_host.to(device, non_blocking=True)])
            host = [t.cpu() for t in r]
        e1.record()
        torch.cuda.synchronize()
        ms_e2e_sync = e0.elapsed_time(e1) / steps
        # the same with two batches in flight (HostDetectionPipeline: per-slot stream, device score buffer, pinned results)
        pipe = rn.pipeline.HostDetectionPipeline(head, B, HW, depth=2)

        def pipe_batches(n):
            prev = None
            for _ in range(n):
                k = pipe.submit(reg_host, cls_host)
                if prev is not None:
                    pipe.result(prev)
                prev = k
            return pipe.result(prev)
        got = pipe_batches(3)
        assert torch.equal(got[0], res[0].cpu()) and torch.equal(got[1], res[1].cpu())
        torch.cuda.synchronize()
        e0.record()
        pipe_batches(steps)
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / steps
        del pipe
        t = torch.tensor([ms, ms_e2e, ms_2s, ms_e2e_copy, ms_e2e_sync], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_2s, ms_e2e_copy, ms_e2e_sync = [float(x) for x in t.cpu()]
        ndet = int((res[1] >= 0).sum().item())
        out[tag] = {"pages_per_s": world * B / (ms * 1e-3), "ms_per_batch": ms,
                    "pages_per_s_two_streams": world * B / (ms_2s * 1e-3),
                    "e2e_pages_per_s": world * B / (ms_e2e * 1e-3), "e2e_api": "HostDetectionPipeline.submit / result, 2 batches in flight",
                    "e2e_one_batch_at_a_time_pages_per_s": world * B / (ms_e2e_sync * 1e-3),
                    "e2e_full_copy_pages_per_s": world * B / (ms_e2e_copy * 1e-3),
                    "detections_per_page": ndet / B}
    out["workload"] = "configs[2]: %d pages/GPU of 800x1333, 1 class, thr 0.05, NMS 0.5, 300 detections" % B
    out["candidates_per_page"] = float((cls_np > 0.05).sum()) / B
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--contiguous-shards", action="store_true", help="several ranks: rank r takes pages [16r, 16r+16) instead of the load-aware deal")
    ap.add_argument("--schedule", default="overlapped", choices=["overlapped", "in-order"],
                    help="what `value` times: K1 of batch s+1 beside K2 of batch s (default) or K1(s) then K2(s); both are measured")
    ap.add_argument("--pipelined", action="store_true", help="several ranks: enqueue K1 + publish of batch s+1 ahead of K2 of batch s "
                    "(TargetLossStep.run_pipelined; measured 3 %% slower than in order at N = 2 and 8, so not the default)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-input leg (used for the ncu launch list of the `value` region)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
