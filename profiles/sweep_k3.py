"""Time the stages of the inference tail (K3 alone, NMS alone, merge alone, the whole graph) at the benchmark's sizes for every
A/B build (librn_b200.<name>.so, see build.py --variant), each stage as a train of back-to-back launches over six input sets in
rotation (rn_debug_filter_stages, the method of bench.py), and print a checksum of the detections so that variants can be
checked against each other bit for bit; one subprocess per build.

    python profiles/sweep_k3.py            # driver
    python profiles/sweep_k3.py --child    # one measurement with the current environment
"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import retinanet_b200 as rn
    import synthetic
    lib = rn._lib.load()
    HW, B, SETS = (800, 1333), 64, 6
    anchors = np.asarray(rn.anchors_for_shape(HW + (3,)))
    _, anns = synthetic.training_batch(3, batch=B)
    cls_np, reg_np = synthetic.inference_predictions(3, B, anchors, anns, classes=1)
    cls, reg = torch.from_numpy(cls_np).cuda(), torch.from_numpy(reg_np).cuda()
    dets = [rn.pipeline.DetectionStep(HW, B, 1) for _ in range(SETS)]
    for k, d in enumerate(dets):
        d.load_predictions(torch.roll(cls, k, 0), torch.roll(reg, k, 0))
        d.run()
    torch.cuda.synchronize()
    busy = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")

    def train(reps=120):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            for _ in range(8):
                busy.amax()
            e0.record()
            for i in range(reps):
                dets[i % SETS]._graph.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        return best

    whole = train()
    h = hashlib.sha1(dets[0].boxes.cpu().numpy().tobytes() + dets[0].scores.cpu().numpy().tobytes()).hexdigest()[:12]
    out = []
    for mask in (1, 2, 4):
        lib.rn_debug_filter_stages(mask)
        for d in dets:
            d._graph = None
            d.run()
        out.append(train())
    lib.rn_debug_filter_stages(7)
    print("K3 %.2f us  NMS %.2f us  merge %.2f us  graph %.2f us  %s" % (out[0], out[1], out[2], whole, h))


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        libdir = os.path.join(ROOT, "retinanet-for-table-detection_b200")
        variants = [""] + sorted(f for f in os.listdir(libdir) if f.startswith("librn_b200.") and f != "librn_b200.so"
                                 and f.endswith(".so"))
        for rep in range(int(os.environ.get("SWEEP_REPS", "2"))):
            for lib in variants:
                env = dict(os.environ)
                if lib:
                    env["RN_B200_LIB"] = os.path.join(libdir, lib)
                out = subprocess.run([sys.executable, __file__, "--child"], env=env, stdout=subprocess.PIPE,
                                     stderr=subprocess.STDOUT, text=True).stdout.strip().splitlines()
                print("%-28s %s" % (lib or "librn_b200.so", out[-1] if out else "?"))
                sys.stdout.flush()
