#!/usr/bin/env python
"""BASELINE config 4 on one GPU: z1/z2 posterior extraction (encoders only) + per-utterance mu2 over synthetic
utterances packed in HBM (lengths U[200,1600] frames, 20-frame segments at stride 8).  Prints segments/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pytorch_scalablefhvae_b200 as P

U = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = P.FHVAE(20 * 80, [256, 256], [256, 256], 32, 32, [256, 256], seg_len=20, num_seqs=1000,
            gemm_mode=P.MODE_BF16X3, use_cuda_graphs=True).to(dev)
lens = np.random.default_rng(7).integers(200, 1601, size=U)
feats = torch.randn(int(lens.sum()), 80, device=dev)
for bs in (256, 2048):
    out = P.extract_posteriors(m, feats[:int(lens[:50].sum())], lens[:50], batch_size=bs)     # warm-up / capture
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = P.extract_posteriors(m, feats, lens, batch_size=bs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    S = out["z1_mu"].shape[0]
    print(f"batch {bs}: {U} utterances, {S} segments in {dt * 1e3:.1f} ms -> {S / dt:,.0f} segments/s, "
          f"{U / dt:,.0f} utterances/s; mu2 {tuple(out['mu2'].shape)}")
