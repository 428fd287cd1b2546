import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import pytorch_scalablefhvae_b200 as P
c = bench.CFG
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = P.FHVAE(c["T"] * c["F"], [c["H"]] * c["L"], [c["H"]] * c["L"], c["Z"], c["Z"], [c["H"]] * c["L"], seg_len=c["T"], num_seqs=c["N"], gemm_mode=P.MODE_BF16X3, use_cuda_graphs=True).to(dev)
opt = P.FusedAdam(m.parameters(), lr=1e-3, betas=(0.95, 0.999))
x, idx, nsegs = bench.synth(c["B"], c["T"], c["F"], c["N"], 1234)
xh, idh, nsh = x.pin_memory(), idx.pin_memory(), nsegs.pin_memory()
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
def step(measure):
    t = time.perf_counter()
    opt.zero_grad(); t = tick("zero_grad", t) if measure else t
    xd = xh.to(dev, non_blocking=True); t = tick("h2d", t) if measure else t
    out = m(xd, idh, c["N"], nsh); t = tick("forward", t) if measure else t
    lss = P.loss_function(out[0], out[1], c["alpha"]); t = tick("loss", t) if measure else t
    lss.backward(); t = tick("backward", t) if measure else t
    opt.step(); t = tick("opt.step", t) if measure else t
    v = float(lss.detach()); t = tick("item(sync)", t) if measure else t
for _ in range(10): step(False)
torch.cuda.synchronize()
n = 100
t0 = time.perf_counter()
for _ in range(n): step(True)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("per step ms", tot / n * 1e3)
for k, v in acc.items(): print(f"  {k:12s} {v / n * 1e6:8.1f} us")

# ---- finer: where inside forward() does the host time go, and when does the GPU get its first work?
import functools
sub = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    @functools.wraps(f)
    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            sub[label] = sub.get(label, 0.0) + time.perf_counter() - t0
    setattr(obj, name, g)
plan = list(m._plans.values())[0]
wrap(m, "_check_ids", "forward._check_ids")
wrap(plan, "load_inputs", "forward.load_inputs")
wrap(plan, "run_forward", "forward.run_forward(graph launch)")
wrap(m, "_publish", "forward._publish")
wrap(plan, "run_backward", "backward.run_backward")
wrap(m, "_deliver_grads", "backward._deliver_grads")
acc.clear()
for _ in range(n): step(True)
torch.cuda.synchronize()
for k, v in sub.items(): print(f"  {k:36s} {v / n * 1e6:8.1f} us")
