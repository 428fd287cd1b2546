"""Debug aid: dump the intermediate buffers of one eager forward (config 1) so that two runs with different FHVAE_PDL
masks can be compared buffer by buffer:  FHVAE_PDL=1 python tools/pdl_debug.py a.pt MODE; FHVAE_PDL=0 ... b.pt; ... cmp a.pt b.pt"""
import sys, torch
sys.path.insert(0, ".")
if sys.argv[1] == "cmp":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for k in a:
        d = (a[k].float() - b[k].float()).abs().max().item()
        print(f"{k:28s} max|a-b| = {d:.3e}   max|b| = {b[k].float().abs().max().item():.3e}")
        if d > 0 and ".h(" in k:
            df = (a[k].float() - b[k].float()).abs().reshape(20, 8, 32, 8, 32)     # [t][group][row][rank][unit]
            print("   per t:", [f"{v:.1e}" for v in df.amax(dim=(1, 2, 3, 4)).tolist()])
            print("   per group:", [f"{v:.1e}" for v in df.amax(dim=(0, 2, 3, 4)).tolist()])
            print("   per rank:", [f"{v:.1e}" for v in df.amax(dim=(0, 1, 2, 4)).tolist()])
            print("   per row (group of first bad):", [f"{v:.0e}" for v in df.amax(dim=(0, 3, 4))[int(df.amax(dim=(0, 2, 3, 4)).argmax())].tolist()])
    sys.exit(0)
import pytorch_scalablefhvae_b200 as P
import numpy as np
def synth_batch(B, T, F, N, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, F, generator=g)
    idx = torch.from_numpy(np.random.default_rng(seed).choice(N, size=B)).long()
    return x, idx, torch.full((B,), 50, dtype=torch.long)
mode = int(sys.argv[2])
graphs = len(sys.argv) > 3 and sys.argv[3] == "graphs"
torch.manual_seed(0)
m = P.FHVAE(1600, [256, 256], [256, 256], 32, 32, [256, 256], seg_len=20, num_seqs=1000, gemm_mode=mode, use_cuda_graphs=graphs).cuda()
x, idx, nsegs = synth_batch(256, 20, 80, 1000)
g = torch.Generator().manual_seed(2)
eps = {"z1": torch.randn(256, 32, generator=g).cuda(), "z2": torch.randn(256, 32, generator=g).cuda()}
out = {}
for rep in range(2):
    res = m(x.cuda(), idx, 1000, nsegs, eps=eps)
    torch.cuda.synchronize()
    pl = list(m._plans.values())[0]
    for name in ("P", "Q", "h", "c", "acts"):
        for k, v in getattr(pl, name).items():
            out[f"r{rep}.{name}{k}"] = v.detach().cpu().clone()
    for name in ("z2head", "z1head", "xhead", "zcat", "bsum"):
        out[f"r{rep}.{name}"] = getattr(pl, name).detach().cpu().clone()
    for k, v in pl.wave_packed.items():
        out[f"r{rep}.packed[{k}]"] = v.detach().cpu().clone()
torch.save(out, sys.argv[1])
