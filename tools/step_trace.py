#!/usr/bin/env python
"""Kernel timeline of the CUDA-graph-replayed train step (CUPTI through torch.profiler): start / end of
every kernel on its stream, so overlap of the side-stream weight-gradient GEMMs with the recurrence
and the gaps on the critical path are visible.  Not a bench (profiler attached)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import pytorch_scalablefhvae_b200 as P  # noqa: E402


def main():
    c = bench.CFG
    dev = torch.device("cuda", 0)
    mode = {"f32": P.MODE_F32_SIMT, "bf16x3": P.MODE_BF16X3, "bf16": P.MODE_BF16}[os.environ.get("FHVAE_MODE", "bf16x3")]
    torch.manual_seed(0)
    m = P.FHVAE(c["T"] * c["F"], [c["H"]] * c["L"], [c["H"]] * c["L"], c["Z"], c["Z"], [c["H"]] * c["L"],
                seg_len=c["T"], num_seqs=c["N"], gemm_mode=mode, use_cuda_graphs=True).to(dev)
    opt = P.FusedAdam(m.parameters(), lr=1e-3, betas=(0.95, 0.999))
    x, idx, nsegs = bench.synth(c["B"], c["T"], c["F"], c["N"], 1234)
    xd, idd, nsd = x.to(dev), idx.to(dev), nsegs.to(dev)
    e2e = "--e2e" in sys.argv
    xh, idh, nsh = x.pin_memory(), idx.pin_memory(), nsegs.pin_memory()

    def step():
        if not e2e:
            return m.train_step(xd, idd, nsd, opt, c["alpha"])
        opt.zero_grad()
        out = m(xh.to(dev, non_blocking=True), idh, c["N"], nsh)
        lss = P.loss_function(out[0], out[1], c["alpha"])
        lss.backward()
        opt.step()
        return float(lss.detach())

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # last full step = everything after the second-to-last Adam kernel up to (and including) the last one
    adams = [i for i, e in enumerate(evs) if "adam_flat" in e.name and "Optimizer" not in e.name]
    if len(adams) < 2:
        print("no complete step recorded", len(evs))
        return
    i0 = adams[-2] + 1
    evs = evs[:adams[-1] + 1]
    t0 = evs[i0].time_range.start
    prev_end = max((e.time_range.end for e in evs[:i0]), default=t0)
    print(f"# gap before step (prev kernel end -> first kernel of the step): {t0 - prev_end:.1f} us")
    print("# start_us  dur_us  stream  name")
    last_end = t0
    for e in evs[i0:]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        nm = e.name
        nm = nm[:nm.index("(")] if "(" in nm else nm
        nm = nm.replace("fhvae::", "").replace("void ", "")[:60]
        print(f"{s:9.1f} {d:8.1f}  {getattr(e, 'stream', '?')}  {nm}")
        last_end = max(last_end, e.time_range.end)
    print(f"# step span {last_end - t0:.1f} us")


if __name__ == "__main__":
    main()
