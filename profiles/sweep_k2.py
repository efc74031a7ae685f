"""Time K2 (and K1) at the benchmark's sizes for a grid of tuning knobs; one subprocess per setting.

    python profiles/sweep_k2.py            # driver: prints one line per setting
    python profiles/sweep_k2.py --child    # one measurement with the current environment
"""
import itertools
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
    HW, B = (800, 1333), 16
    anchors = rn.anchors_for_shape(HW + (3,))
    N = anchors.shape[0]
    images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors))
    cls, reg = synthetic.training_predictions(2, B, N, classes=1)
    step = rn.pipeline.TargetLossStep(HW + (3,), B, 22, 1)
    step.load_annotations(images, anns)
    step.load_predictions(torch.from_numpy(cls), torch.from_numpy(reg))
    for _ in range(5):
        step.run()
    steps = 100
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        step.run(events=evs[i])
    torch.cuda.synchronize()
    k1 = sorted(e[0].elapsed_time(e[1]) for e in evs)
    k2 = sorted(e[1].elapsed_time(e[2]) for e in evs)
    l = step.losses.cpu().numpy()
    gsum = float(step.grad_cls.double().abs().sum() + step.grad_reg.double().abs().sum())
    print("K1 %.2f us  K2 mean %.2f med %.2f min %.2f us  losses %.7f %.7f  |grad| %.9f"
          % (1e3 * sum(k1) / steps, 1e3 * sum(k2) / steps, 1e3 * k2[steps // 2], 1e3 * k2[0], l[0], l[1], gsum))


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        settings = [{"RN_K2_FAST": "0"}]
        libdir = os.path.join(ROOT, "retinanet-for-table-detection_b200")
        variants = [""] + sorted(f for f in os.listdir(libdir) if f.startswith("librn_b200.") and f != "librn_b200.so"
                                 and f.endswith(".so"))
        for lib, (up, minb) in itertools.product(variants, ((1, 4), (1, 6), (1, 8), (2, 4), (2, 6), (2, 8), (4, 4), (4, 6))):
            st = {"RN_K2_UP": str(up), "RN_K2_MINB": str(minb)}
            if lib:
                st["RN_B200_LIB"] = os.path.join(libdir, lib)
            settings.append(st)
        for st in settings:
            env = dict(os.environ, **st)
            out = subprocess.run([sys.executable, __file__, "--child"], env=env, stdout=subprocess.PIPE,
                                 stderr=subprocess.STDOUT, text=True).stdout.strip().splitlines()
            print(" ".join("%s=%s" % (k, os.path.basename(v)) for k, v in st.items()), "|", out[-1] if out else "?")
            sys.stdout.flush()
