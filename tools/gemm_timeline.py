"""Phase timeline of the CTA-pair GEMM kernel (clock64 stamps) + device-time sweeps (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytorch_distributed_nlp_b200 import _lib as L
dev = torch.device("cuda", 0)
bf = torch.bfloat16


def args(M, N, K, A, B, D, am=0, bm=0, epi=0, bias=None, aux_in=None, aux_out=None, kernel=2, bn=256, timing=None, ws=None, splits=0):
    a = L.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_major = A.data_ptr(), A.shape[1], am
    a.B, a.ldb, a.b_major = B.data_ptr(), B.shape[1], bm
    a.D, a.ldd, a.epilogue = D.data_ptr(), D.shape[1], epi
    a.bias = L.ptr(bias); a.aux_in = L.ptr(aux_in); a.ld_aux_in = aux_in.shape[1] if aux_in is not None else 0
    a.aux_out = L.ptr(aux_out); a.ld_aux_out = aux_out.shape[1] if aux_out is not None else 0
    a.force_kernel, a.force_bn, a.force_splits = kernel, bn, splits
    a.debug_timing = L.ptr(timing)
    a.workspace, a.workspace_bytes = L.ptr(ws), (ws.numel() if ws is not None else 0)
    return a


def device_time(a, reps=20):
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        L.call("b2_gemm_bf16", a, s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            L.call("b2_gemm_bf16", a, torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def timeline(M, N, K, epi=0, label=""):
    A = torch.randn(M, K, device=dev).to(bf); B = (torch.randn(N, K, device=dev) * .05).to(bf)
    D = torch.empty(M, N, dtype=bf, device=dev); bias = torch.randn(N, device=dev).to(bf)
    U = torch.empty(M, N, dtype=bf, device=dev)
    kw = {}
    if epi == L.EPI_BIAS_GELU: kw = dict(bias=bias, aux_out=U)
    if epi == L.EPI_BIAS: kw = dict(bias=bias)
    t_us = device_time(args(M, N, K, A, B, D, epi=epi, **kw))
    grid = 2 * min(74, ((M + 255) // 256) * (N // 256))
    timing = torch.zeros(grid, 8, dtype=torch.int64, device=dev)
    L.call("b2_gemm_bf16", args(M, N, K, A, B, D, epi=epi, timing=timing, **kw), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t = timing.cpu()
    lead = t[0::2]
    base = lead[:, 0:1]
    rel = (lead - base).float()
    med = rel.median(0).values.tolist()
    tiles = ((M + 255) // 256) * (N // 256)
    print("%-28s M%d N%d K%d tiles/pair %.2f  device %.1f us (%.0f TF/s) | cycles since entry (median over leader CTAs): "
          "setup %d, first-data %d, last-mma-issued %d, last-acc-done %d, epilogue-done %d, exit %d" % (
              label, M, N, K, tiles / min(74, tiles), t_us, 2.0 * M * N * K / t_us / 1e6, med[1], med[2], med[3], med[4], med[5], med[6]), flush=True)


if __name__ == "__main__":
    for K in (768, 1536, 3072, 6144):
        timeline(4096, 768, K, label="N768 sweep K (EPI_NONE)")
    for N in (2304, 3072):
        timeline(4096, N, 768, label="K768 multi-tile (EPI_NONE)")
    timeline(4096, 3072, 768, epi=L.EPI_BIAS_GELU, label="ffn1 (BIAS_GELU)")
    timeline(4096, 2304, 768, epi=L.EPI_BIAS, label="qkv (BIAS)")
    timeline(8192, 4096, 4096, label="big square-ish (EPI_NONE)")
    # single-CTA kernel for comparison (device time only)
    for (M, N, K) in ((4096, 768, 3072), (4096, 3072, 768), (8192, 4096, 4096)):
        A = torch.randn(M, K, device=dev).to(bf); B = (torch.randn(N, K, device=dev) * .05).to(bf)
        D = torch.empty(M, N, dtype=bf, device=dev)
        t_us = device_time(args(M, N, K, A, B, D, kernel=1, bn=256))
        print("single-CTA 128x256  M%d N%d K%d  device %.1f us (%.0f TF/s)" % (M, N, K, t_us, 2.0 * M * N * K / t_us / 1e6))
        t_ref = None
        C = torch.empty(M, N, dtype=bf, device=dev)
        for _ in range(3): torch.matmul(A, B.t(), out=C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): torch.matmul(A, B.t(), out=C)
        e1.record(); torch.cuda.synchronize()
        t_ref = e0.elapsed_time(e1) * 1e3 / 20
        print("cuBLAS (torch.matmul) same shape      device %.1f us (%.0f TF/s)  [comparison baseline only]" % (t_ref, 2.0 * M * N * K / t_ref / 1e6))
