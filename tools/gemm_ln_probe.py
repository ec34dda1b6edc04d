"""b2_gemm_ln_fwd (cluster kernel) vs the GEMM + LayerNorm pair it replaces, timed as captured loops on one B200:
how many 8-CTA clusters the device co-schedules, and how the fused kernel's time moves with the number of row blocks
(= clusters) -- a jump between two cluster counts is a second wave."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytorch_distributed_nlp_b200 import _lib as L

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = L.load()
bf = torch.bfloat16
for H in (768, 1024):
    print("hidden %d: max co-resident clusters %d" % (H, lib.b2_gemm_ln_max_clusters(H)), flush=True)


def args(M, N, K, A, B, D, bias, X, rng, p):
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_major = A.data_ptr(), K, L.MAJOR_K
    a.B, a.ldb, a.b_major = B.data_ptr(), K, L.MAJOR_K
    a.D, a.ldd, a.epilogue = D.data_ptr(), N, L.EPI_BIAS_DROPOUT_RESIDUAL
    a.bias, a.aux_in, a.ld_aux_in = bias.data_ptr(), X.data_ptr(), N
    a.dropout_p, a.rng_state, a.rng_site = p, rng.data_ptr(), 5
    return a


def timed(fn, sets, rep=4):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            for s in sets:
                fn(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rep * len(sets))


def probe(M, N, K, p=0.1):
    nset = max(4, min(24, int(300e6 // (2 * M * K + 2 * N * K + 6 * M * N)) + 1))
    sets = []
    for _ in range(nset):
        sets.append(dict(A=torch.randn(M, K, device=dev).to(bf), B=(torch.randn(N, K, device=dev) * 0.05).to(bf),
                         bias=torch.randn(N, device=dev).to(bf), X=torch.randn(M, N, device=dev).to(bf),
                         Xf=torch.randn(M, N, device=dev), Yf=torch.empty(M, N, device=dev),
                         Z=torch.empty(M, N, dtype=bf, device=dev), Y=torch.empty(M, N, dtype=bf, device=dev),
                         mean=torch.empty(M, device=dev), rstd=torch.empty(M, device=dev)))
    gamma, beta = torch.ones(N, device=dev).to(bf), torch.zeros(N, device=dev).to(bf)
    rng = torch.tensor([77, 3], dtype=torch.int64, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream

    def fused(s):
        L.call("b2_gemm_ln_fwd", args(M, N, K, s["A"], s["B"], s["Z"], s["bias"], s["Xf"], rng, p), gamma.data_ptr(),
               beta.data_ptr(), 1e-12, s["Y"].data_ptr(), N, s["Yf"].data_ptr(), N, s["mean"].data_ptr(),
               s["rstd"].data_ptr(), st())

    def gemm_only(s):
        L.call("b2_gemm_bf16", args(M, N, K, s["A"], s["B"], s["Z"], s["bias"], s["X"], rng, p), st())

    def pair(s):
        gemm_only(s)
        L.call("b2_layernorm_fwd", s["Z"].data_ptr(), gamma.data_ptr(), beta.data_ptr(), M, N, 1e-12, s["Y"].data_ptr(),
               s["mean"].data_ptr(), s["rstd"].data_ptr(), st())

    tf, tg, tp = timed(fused, sets), timed(gemm_only, sets), timed(pair, sets)
    fl = 2.0 * M * N * K
    print("M %5d N %4d K %4d  clusters %3d | fused %6.2f us (%6.1f TF/s) | gemm %6.2f us | gemm+LN %6.2f us (%6.1f TF/s)"
          % (M, N, K, (M + 255) // 256, tf, fl / tf / 1e6, tg, tp, fl / tp / 1e6), flush=True)


def timeline(M, N, K, p=0.1):
    """clock64 stamps of the first epilogue warp of every CTA (b2_gemm_args_t.debug_timing), as phase durations"""
    A, B = torch.randn(M, K, device=dev).to(bf), (torch.randn(N, K, device=dev) * 0.05).to(bf)
    bias, Xf = torch.randn(N, device=dev).to(bf), torch.randn(M, N, device=dev)
    Z, Y = torch.empty(M, N, dtype=bf, device=dev), torch.empty(M, N, dtype=bf, device=dev)
    Yf, mean, rstd = torch.empty(M, N, device=dev), torch.empty(M, device=dev), torch.empty(M, device=dev)
    gamma, beta = torch.ones(N, device=dev).to(bf), torch.zeros(N, device=dev).to(bf)
    rng = torch.tensor([77, 3], dtype=torch.int64, device=dev)
    ctas = ((M + 255) // 256) * 2 * (N // 256)
    stamps = torch.zeros(ctas, 8, dtype=torch.int64, device=dev)
    a = args(M, N, K, A, B, Z, bias, Xf, rng, p)
    for it in range(3):
        a.debug_timing = stamps.data_ptr() if it == 2 else None
        L.call("b2_gemm_ln_fwd", a, gamma.data_ptr(), beta.data_ptr(), 1e-12, Y.data_ptr(), N, Yf.data_ptr(), N,
               mean.data_ptr(), rstd.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    t = stamps.cpu().double()
    d = t[:, 1:] - t[:, :-1]
    names = ["entry -> set up", "-> accumulator complete", "-> pass 1 done", "-> statistics complete", "-> pass 2 done",
             "-> tiles read out", "-> cluster drained"]
    print("M %d N %d K %d: %d CTAs, SM cycles (mean / max over CTAs)" % (M, N, K, ctas))
    for i, n in enumerate(names):
        print("   %-26s %8.0f %8.0f" % (n, float(d[:, i].mean()), float(d[:, i].max())))
    print("   %-26s %8.0f %8.0f" % ("entry -> drained", float((t[:, 7] - t[:, 0]).mean()), float((t[:, 7] - t[:, 0]).max())),
          flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "--timeline":
    for spec in sys.argv[2:]:
        M, N, K = (int(v) for v in spec.split("x"))
        timeline(M, N, K)
    sys.exit(0)
if len(sys.argv) > 1:          # explicit shapes: MxNxK ...
    for spec in sys.argv[1:]:
        M, N, K = (int(v) for v in spec.split("x"))
        probe(M, N, K)
    sys.exit(0)
for K in (768, 3072):
    for M in (256, 512, 1024, 2048, 3072, 3584, 4096, 4352, 8192):
        probe(M, 768, K)
for K in (1024, 4096):
    for M in (1024, 2048, 4096):
        probe(M, 1024, K)
probe(4096, 768, 3072, p=0.0)
