"""Diagnostic run of the two tcgen05 kernels against torch on raw bf16 operands.
Prints per-shape errors (and a corner of the result when wrong); exit code 1 on mismatch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posterior_matching_b200 import _lib

torch.manual_seed(0)
S = torch.cuda.current_stream().cuda_stream
bad = 0


def report(tag, got, want):
    global bad
    err = (got.double() - want.double()).abs().max().item()
    scale = want.double().abs().max().item() + 1e-30
    ok = err <= 2e-3 * scale and torch.isfinite(got).all().item()
    print(f"{'OK ' if ok else 'BAD'} {tag}: max abs err {err:.3e} (scale {scale:.3e})", flush=True)
    if not ok:
        bad += 1
        print("  got :", got[:4, :8].tolist())
        print("  want:", want[:4, :8].tolist())
        nz = (got != 0).float().mean().item()
        print(f"  nonzero frac {nz:.3f}; row-err profile (first 16 rows):",
              (got.double() - want.double()).abs().max(1).values[:16].tolist())
        print("  col-err profile (first 16 cols):", (got.double() - want.double()).abs().max(0).values[:16].tolist())


def nt(M, N, K, lda=None, ldb=None, bias=True):
    lda = lda or K
    ldb = ldb or K
    A = torch.zeros(M, lda, device="cuda", dtype=torch.bfloat16)
    Bt = torch.zeros(N, ldb, device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Bt[:, :K] = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bv = torch.randn(N, device="cuda") if bias else None
    y = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(_lib.lib.pmvae_tc_gemm_nt(A.data_ptr(), lda, Bt.data_ptr(), ldb, bv.data_ptr() if bias else None, M, N, K,
                                         y.data_ptr(), S), "nt")
    torch.cuda.synchronize()
    want = A[:, :K].double() @ Bt[:, :K].double().T + (bv.double() if bias else 0)
    report(f"nt M={M} N={N} K={K} lda={lda} ldb={ldb}", y, want)


def tn(M, N, rows, lda=None, ldb=None):
    lda = lda or M
    ldb = ldb or N
    A = torch.zeros(rows, lda, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(rows, ldb, device="cuda", dtype=torch.bfloat16)
    A[:, :M] = torch.randn(rows, M, device="cuda").to(torch.bfloat16)
    B[:, :N] = (torch.randn(rows, N, device="cuda") / rows ** 0.5).to(torch.bfloat16)
    y = torch.zeros(M, N, device="cuda")
    _lib.check(_lib.lib.pmvae_tc_gemm_tn(A.data_ptr(), lda, B.data_ptr(), ldb, M, N, rows, y.data_ptr(), S), "tn")
    torch.cuda.synchronize()
    want = A[:, :M].double().T @ B[:, :N].double()
    report(f"tn M={M} N={N} rows={rows} lda={lda} ldb={ldb}", y, want)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "nt"):
    for shp in [(128, 256, 64), (128, 256, 256), (256, 256, 256), (1000, 256, 256), (4096, 152, 256), (300, 256, 152),
                (77, 8, 256), (640, 2144, 256), (513, 256, 2144), (200, 256, 8), (70000, 256, 256)]:
        nt(*shp)
if which in ("all", "tn"):
    for shp in [(256, 256, 64), (256, 256, 128), (256, 256, 4096), (256, 256, 1000), (256, 152, 512), (256, 8, 300),
                (256, 2144, 777), (256, 256, 70000)]:
        tn(*shp)
print("bad =", bad)
sys.exit(1 if bad else 0)
