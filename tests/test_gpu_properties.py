"""GPU: hypothesis property tests of the exact index kernels: mu2 gather, deterministic
scatter-reduce ("rows touched" set), MAP accumulate and sparse row copies -- N = 1 tables, B = 1 batches,
all-duplicate batches, first/last rows, any Z multiple of 4.  Integers in fp32 make every sum exact, so the
comparisons are bit-exact."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from pytorch_scalablefhvae_b200.plan import ptr
from util import call

pytestmark = pytest.mark.gpu
DEV = "cuda"
SET = settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@st.composite
def table_case(draw):
    N = draw(st.sampled_from([1, 2, 3, 17, 1000]))
    Z = draw(st.sampled_from([4, 8, 16, 32, 64, 132]))
    B = draw(st.sampled_from([1, 2, 31, 32, 33, 257]))
    kind = draw(st.sampled_from(["random", "all_same", "first_last", "sorted"]))
    seed = draw(st.integers(0, 2 ** 16))
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, N, (B,), generator=g)
    if kind == "all_same":
        idx[:] = int(idx[0])
    elif kind == "first_last":
        idx[::2] = 0
        idx[1::2] = N - 1
    elif kind == "sorted":
        idx = idx.sort().values
    return N, Z, B, idx, seed


@SET
@given(table_case())
def test_gather_property(case):
    N, Z, B, idx, seed = case
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(N, Z, generator=g).to(DEV)
    out = torch.full((B, Z), -1.0, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    idd = idx.to(DEV)                      # (device temporaries must outlive the launch: keep references)
    call("fhvae_mu2_gather", ptr(table), ptr(idd), ptr(out), B, Z, N, ptr(flag))
    assert torch.equal(out, table[idd]) and int(flag) == 0


@SET
@given(table_case())
def test_scatter_reduce_property(case):
    """dst[idx[b]] += src[b] summed over duplicates == index_add_; touched == first occurrence of each distinct row;
    rows not in idx stay bit-identical; two runs are bit-identical (no float atomics)."""
    N, Z, B, idx, seed = case
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(-8, 9, (B, Z), generator=g).float()
    base = torch.randint(-3, 4, (N, Z), generator=g).float()
    ref = base.clone().index_add_(0, idx, src)
    outs = []
    sd, idd = src.to(DEV), idx.to(DEV)
    for _ in range(2):
        dst = base.clone().to(DEV)
        touched = torch.full((B,), -1, dtype=torch.int32, device=DEV)
        call("fhvae_mu2_scatter_reduce", ptr(sd), ptr(idd), ptr(dst), ptr(touched), B, Z, N)
        outs.append(dst.cpu())
    assert torch.equal(outs[0], ref) and torch.equal(outs[0], outs[1])
    first, seen = torch.zeros(B, dtype=torch.int32), set()
    for b, r in enumerate(idx.tolist()):
        if r not in seen:
            first[b] = 1
            seen.add(r)
    assert torch.equal(touched.cpu(), first)
    assert int(first.sum()) == len(set(idx.tolist()))


@SET
@given(table_case())
def test_accumulate_property(case):
    """utils.py:49-56 batched: zsum[k] = sum of rows with label k, cnt[k] = their number (exact)."""
    N, Z, B, idx, seed = case
    g = torch.Generator().manual_seed(seed)
    z = torch.randint(-8, 9, (B, 2 * Z), generator=g).float()        # (B, 2Z) head layout: ld = 2Z, first Z columns used
    zsum, cnt = torch.zeros(N, Z, device=DEV), torch.zeros(N, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    zd, idd = z.to(DEV), idx.to(DEV)
    call("fhvae_mu2_accumulate", ptr(zd), 2 * Z, ptr(idd), ptr(zsum), ptr(cnt), B, Z, N, ptr(flag))
    assert torch.equal(zsum.cpu(), torch.zeros(N, Z).index_add_(0, idx, z[:, :Z]))
    assert torch.equal(cnt.cpu().long(), torch.bincount(idx, minlength=N)) and int(flag) == 0


@SET
@given(st.integers(1, 64), st.sampled_from([4, 16, 32]), st.integers(0, 2 ** 16))
def test_rows_copy_property(n, Z, seed):
    """dst[dst_rows[i]] = src[src_rows[i]] for a permutation of destination rows (no write conflicts); negative rows
    are skipped; untouched destination rows keep their bits."""
    g = torch.Generator().manual_seed(seed)
    S, D = n + 5, n + 3
    src = torch.randn(S, Z, generator=g)
    dst0 = torch.randn(D, Z, generator=g)
    s_rows = torch.randint(0, S, (n,), generator=g)
    d_rows = torch.randperm(D, generator=g)[:n]
    skip = torch.rand(n, generator=g) < 0.2
    d_rows = torch.where(skip, torch.full_like(d_rows, -1), d_rows)
    dst = dst0.clone().to(DEV)
    sd, srd, drd = src.to(DEV), s_rows.to(DEV), d_rows.to(DEV)
    call("fhvae_rows_copy", ptr(sd), ptr(srd), ptr(dst), ptr(drd), n, Z)
    ref = dst0.clone()
    for i in range(n):
        if int(d_rows[i]) >= 0:
            ref[int(d_rows[i])] = src[int(s_rows[i])]
    assert torch.equal(dst.cpu(), ref)


def test_out_of_range_rows_never_touch_memory():
    """gather / scatter / accumulate with rows outside [0,N): NaN + flag (gather), skipped (others)."""
    N, Z, B = 5, 8, 6
    table = torch.arange(N * Z, dtype=torch.float32, device=DEV).view(N, Z)
    idx = torch.tensor([0, 4, 5, -1, 2, 1 << 40], device=DEV)
    guard = torch.full((3 * N, Z), 7.0, device=DEV)           # the table sits in the middle of a guarded allocation
    guard[N:2 * N] = table
    tab = guard[N:2 * N]
    out = torch.zeros(B, Z, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    call("fhvae_mu2_gather", ptr(tab), ptr(idx), ptr(out), B, Z, N, ptr(flag))
    assert int(flag) == 2
    ok = torch.tensor([True, True, False, False, True, False], device=DEV)
    assert torch.equal(out[ok], table[idx[ok]]) and bool(torch.isnan(out[~ok]).all())
    src = torch.ones(B, Z, device=DEV)
    touched = torch.zeros(B, dtype=torch.int32, device=DEV)
    call("fhvae_mu2_scatter_reduce", ptr(src), ptr(idx), ptr(tab), ptr(touched), B, Z, N)
    assert float((guard[:N] - 7.0).abs().max()) == 0.0 and float((guard[2 * N:] - 7.0).abs().max()) == 0.0
    assert touched.tolist() == [1, 1, 0, 0, 1, 0]
    zsum, cnt = torch.zeros(N, Z, device=DEV), torch.zeros(N, device=DEV)
    flag.zero_()
    call("fhvae_mu2_accumulate", ptr(src), Z, ptr(idx), ptr(zsum), ptr(cnt), B, Z, N, ptr(flag))
    assert int(flag) == 2 and cnt.tolist() == [1, 0, 1, 0, 1]
