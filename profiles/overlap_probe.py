"""Step time of the three schedules of TargetLossStep at the benchmark's sizes (one GPU): in order (one graph),
pipelined (K1 of the next batch, then K2 of this one), overlapped (the same two concurrently on two streams)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import retinanet_b200 as rn, synthetic
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
torch.cuda.synchronize()
ref = (step.losses.clone(), step.grad_cls.clone(), step.grad_reg.clone())

def timed(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print("in order   %.2f us/step" % timed(step.run))
print("pipelined  %.2f us/step" % timed(step.run_pipelined))
assert torch.equal(step.losses, ref[0]) and torch.equal(step.grad_cls, ref[1]) and torch.equal(step.grad_reg, ref[2])
print("overlapped %.2f us/step" % timed(lambda: step.run_pipelined(overlap=True)))
assert torch.equal(step.losses, ref[0]) and torch.equal(step.grad_cls, ref[1]) and torch.equal(step.grad_reg, ref[2])
print("results identical")
