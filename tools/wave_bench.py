"""A/B micro-benchmark of the wavefront LSTM kernels (development aid).

    python tools/wave_bench.py name=path/to/lib.so [name2=...]   [--modes 1,2] [--timeline]

For every library: one 2-layer stack forward (fhvae_lstm_wave_fwd_planes) and BPTT (fhvae_lstm_wave_bwd_planes) at
the config-1 shape (T=20, B=256, H=256), replayed 10x from a CUDA graph, device time per launch in us; a checksum of
every output is compared with the first library's (variants that only re-time the exchange must be bit-identical),
and the first library is checked against the exact fp32 SIMT recurrence (fhvae_lstm_fwd / fhvae_lstm_bwd, mode 0).
Libraries built with -DFHVAE_TIMELINE also print the in-kernel phase stamps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

T, B, H, L = 20, 256, 256, 2
REPS = 10
STRESS = 0
PACK = True


def vp(t, off=0):
    return ctypes.c_void_p(t.data_ptr() + off * t.element_size()) if t is not None else None


def load(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    for n in ("fhvae_lstm_wave_xchg_bytes", "fhvae_lstm_wave_bwd_xchg_bytes"):
        getattr(lib, n).restype = ctypes.c_longlong
    lib.fhvae_last_error_string.restype = ctypes.c_char_p
    return lib


def graph_us(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(s.cuda_stream)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(REPS):
            fn(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / REPS)
    return best


def run(lib, mode, timeline=False):
    NEWABI = hasattr(lib, "fhvae_lstm_wave_pack")
    g = torch.Generator(device="cuda").manual_seed(0)
    z = lambda *s, sc=0.3: torch.randn(*s, device="cuda", generator=g) * sc
    P, Q = z(T, B, 4 * H), z(B, 4 * H)
    W0, Wi1, W1, b1 = z(4 * H, H, sc=0.05), z(4 * H, H, sc=0.05), z(4 * H, H, sc=0.05), z(4 * H)
    f = lambda *s: torch.zeros(*s, device="cuda")
    h, c, a = [f(T, B, H), f(T, B, H)], [f(T, B, H), f(T, B, H)], [f(T, B, 4 * H), f(T, B, 4 * H)]
    hp = [torch.zeros(2, T * B * H, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    packed = None
    if PACK and hasattr(lib, "fhvae_lstm_wave_pack"):
        lib.fhvae_lstm_wave_pack_bytes.restype = ctypes.c_longlong
        packed = torch.zeros(lib.fhvae_lstm_wave_pack_bytes(H, L, mode) // 4, device="cuda")
        assert lib.fhvae_lstm_wave_pack(vp(W0), vp(Wi1), vp(W1), vp(packed), H, L, mode, None) == 0
        t_pack = graph_us(lambda st: lib.fhvae_lstm_wave_pack(vp(W0), vp(Wi1), vp(W1), vp(packed), H, L, mode, ctypes.c_void_p(st)))
        print(f"     pack kernel {t_pack:.1f} us", flush=True)
    xf = torch.zeros(lib.fhvae_lstm_wave_xchg_bytes(T, B, H, L) // 4, device="cuda")
    xb = torch.zeros(lib.fhvae_lstm_wave_bwd_xchg_bytes(T, B, H, L) // 4, device="cuda")

    def fwd(st):
        r = lib.fhvae_lstm_wave_fwd_planes(vp(P), vp(Q), vp(W0), vp(h[0]), vp(c[0]), vp(a[0]), vp(Wi1), vp(b1), vp(W1),
                                           vp(h[1]), vp(c[1]), vp(a[1]), vp(xf), vp(hp[0]), vp(hp[1]),
                                           ctypes.c_longlong(T * B * H), *([vp(packed)] if NEWABI else []), T, B, H, L, mode,
                                           ctypes.c_void_p(st))
        assert r == 0, lib.fhvae_last_error_string()
    t_f = graph_us(fwd)
    dh_all, dhl1, dhl0 = z(T, B, H), z(B, H), z(B, H)
    dgp = [torch.zeros(2, T * B * 4 * H, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    dgs = [f(B, 4 * H), f(B, 4 * H)]

    def bwd(st):
        r = lib.fhvae_lstm_wave_bwd_planes(vp(dh_all), vp(dhl1), vp(dhl0), vp(W1), vp(c[1]), vp(a[1]), None, vp(dgs[1]),
                                           vp(Wi1), vp(W0), vp(c[0]), vp(a[0]), None, vp(dgs[0]), vp(xb), vp(dgp[1]),
                                           vp(dgp[0]), ctypes.c_longlong(T * B * 4 * H), *([vp(packed)] if NEWABI else []),
                                           T, B, H, L, mode, ctypes.c_void_p(st))
        assert r == 0, lib.fhvae_last_error_string()
    t_b = graph_us(bwd)
    outs = h + c + a + hp + dgp + dgs
    sums = [float(o.double().abs().sum()) for o in outs]
    if STRESS:          # race hunting: every launch must reproduce the first one bit for bit
        st = torch.cuda.current_stream().cuda_stream
        ref_f = [o.clone() for o in h + c + a + hp]
        ref_b = [o.clone() for o in dgp + dgs]
        bad = 0
        for i in range(STRESS):
            fwd(st)
            bwd(st)
            if i % 10 == 9 or i == STRESS - 1:
                bad += sum(int(not torch.equal(x, y)) for x, y in zip(h + c + a + hp, ref_f))
                bad += sum(int(not torch.equal(x, y)) for x, y in zip(dgp + dgs, ref_b))
        print(f"     stress x{STRESS}: {'ok' if bad == 0 else 'MISMATCHES ' + str(bad)}", flush=True)
    tl = None
    if timeline and hasattr(lib, "fhvae_debug_wave_timeline"):
        buf = (ctypes.c_longlong * (2 * 32 * 16))()
        fwd(torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
        lib.fhvae_debug_wave_timeline(buf)
        tl = [[[buf[(layer * 32 + t) * 16 + k] for k in range(16)] for t in range(T + 1)] for layer in range(2)]
    ref = None
    if mode == 1:       # exact fp32 SIMT recurrence of layer 0 (forward h) as the numeric anchor
        h0, c0, a0 = f(T, B, H), f(T, B, H), f(T, B, 4 * H)
        r = lib.fhvae_lstm_fwd(vp(P), vp(Q), vp(W0), vp(h0), vp(c0), vp(a0), vp(f(16, B, H)), T, B, H, 0, None)
        assert r == 0
        ref = float((h0 - h[0]).abs().max() / h0.abs().max())
    return t_f, t_b, sums, tl, ref


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    modes = [1]
    for a in sys.argv[1:]:
        if a.startswith("--modes"):
            modes = [int(v) for v in a.split("=")[1].split(",")]
    timeline = "--timeline" in sys.argv
    global STRESS, PACK
    PACK = "--no-pack" not in sys.argv
    for a in sys.argv[1:]:
        if a.startswith("--stress"):
            STRESS = int(a.split("=")[1])
    base = {}
    for spec in args:
        name, path = spec.split("=")
        lib = load(path)
        for mode in modes:
            t_f, t_b, sums, tl, ref = run(lib, mode, timeline)
            same = "-"
            if mode in base:
                same = "bit-identical" if sums == base[mode] else "DIFFERENT " + str(
                    [i for i, (x, y) in enumerate(zip(sums, base[mode])) if x != y])
            else:
                base[mode] = sums
            print(f"{name:28s} mode {mode}: fwd {t_f:7.1f} us  bwd {t_b:7.1f} us  sum {t_f + t_b:7.1f}   vs first: {same}"
                  + (f"   h0 vs fp32 SIMT {ref:.1e}" if ref is not None else ""), flush=True)
            if tl:
                for layer in range(2):
                    tops = [tl[layer][t][0] for t in range(T + 1)]
                    print(f"     L{layer} periods: " + " ".join(str(tops[t + 1] - tops[t]) for t in range(T - 1 if layer else T))
                          + f" | total {tops[T - 1 if layer else T] - tops[0]}")
                    e = tl[layer]
                    print(f"     L{layer}: entry->prologue done {e[1][15] - e[0][15]}  prologue done->loop end {e[2][15] - e[1][15]}  "
                          f"first top - entry {tops[0] - e[0][15]}")
                    if e[3][15]:
                      print(f"        prologue: alloc+sync {e[3][15] - e[0][15]}  reg_inc {e[4][15] - e[3][15]}  W_hh->TMEM {e[5][15] - e[4][15]}  "
                          f"W_ih1->smem {e[6][15] - e[5][15]}  sync {e[7][15] - e[6][15]}  Q {e[1][15] - e[7][15]}")
                    for t in range(9, 12):
                        r = tl[layer][t]
                        print(f"     L{layer} t{t}: loads_issued {r[6] - r[0]} waited+stored {r[7] - r[6]} pull_total {r[1] - r[0]} "
                              f"rec_done {r[2] - r[1]} tmem_ld {r[3] - r[2]} gates {r[4] - r[3]} cell {r[5] - r[4]} "
                              f"rest {tl[layer][t + 1][0] - r[5]} period {r[0] - tl[layer][t - 1][0]}")
                        if r[13]:
                            n = tl[layer][t + 1]
                            print(f"        rel. to step top: rec_done@{r[2] - r[0]} gates_done@{r[4] - r[0]} cell_done@{r[5] - r[0]} "
                                  f"pub_stored@{r[8] - r[0]} bar@{r[9] - r[0]} released@{r[10] - r[0]} | next step: flag_seen@{n[11] - r[0]} "
                                  f"copies_issued@{n[12] - r[0]} hb_full@{n[13] - r[0]} mma_issued@{n[14] - r[0]} rec_done@{n[2] - r[0]}")


if __name__ == "__main__":
    main()
