"""Time K1 at the benchmark's sizes for every A/B build (librn_b200.<name>.so, see build.py --variant) and print a
checksum of its outputs so that variants can be checked against each other bit for bit; one subprocess per build.

    python profiles/sweep_k1.py            # driver
    python profiles/sweep_k1.py --child    # one measurement with the current environment
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
    out = []
    for cfg, HW, B, gmax in ((2, (800, 1333), 16, 22), (4, (1600, 2400), 4, 22)):
        anchors = rn.anchors_for_shape(HW + (3,))
        images, anns = synthetic.training_batch(cfg, batch=B, anchors=np.asarray(anchors))
        step = rn.pipeline.TargetLossStep(HW + (3,), B, gmax, 1)
        step.load_annotations(images, anns)
        step._build_graphs()
        for _ in range(5):
            step._graphs[0].replay()
        steps = 100
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            step._graphs[0].replay()
        e1.record()
        torch.cuda.synchronize()
        h = hashlib.sha1(step.y_reg.cpu().numpy().tobytes() + step.y_cls.cpu().numpy().tobytes() + step.npos.cpu().numpy().tobytes()).hexdigest()[:12]
        out.append("cfg%d K1 %.2f us %s" % (cfg, 1e3 * e0.elapsed_time(e1) / steps, h))
    print(" | ".join(out))


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
