"""Does the background AdamW really run BESIDE the GEMM CTAs?  A GEMM loop on one stream, the optimizer update of the
whole parameter space on another: alone, alone, together (wall = max over the two streams' events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytorch_distributed_nlp_b200 import _lib as L

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L.load()
bf = torch.bfloat16
M, N, K = 4096, 3072, 768
NSET = 6
sets = [dict(A=torch.randn(M, K, device=dev).to(bf), B=(torch.randn(N, K, device=dev) * 0.05).to(bf),
             D=torch.empty(M, N, dtype=bf, device=dev), bias=torch.randn(N, device=dev).to(bf)) for _ in range(NSET)]
n = 102_272_264 // 8 * 8
master, m, v = torch.randn(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
g = (torch.randn(n, device=dev) * 0.01).to(bf)
shadow = torch.empty(n, dtype=bf, device=dev)
decay = torch.ones(n // 8, dtype=torch.uint8, device=dev)
step = torch.zeros(1, dtype=torch.int64, device=dev)
step_size = torch.zeros(1, device=dev)
hp = L.AdamWHParams()
hp.lr, hp.beta1, hp.beta2, hp.eps, hp.weight_decay, hp.correct_bias = 3e-5, 0.9, 0.999, 1e-6, 0.01, 1
L.call("b2_adamw_prepare", hp, step.data_ptr(), step_size.data_ptr(), torch.cuda.current_stream().cuda_stream)
s_gemm, s_opt = torch.cuda.Stream(priority=-1), torch.cuda.Stream()
NB = 13
chunk = n // NB // 8 * 8


def gemms(stream, reps=60):
    for r in range(reps):
        s = sets[r % NSET]
        a = L.GemmArgs()
        a.M, a.N, a.K = M, N, K
        a.A, a.lda, a.a_major = s["A"].data_ptr(), K, L.MAJOR_K
        a.B, a.ldb, a.b_major = s["B"].data_ptr(), K, L.MAJOR_K
        a.D, a.ldd, a.epilogue = s["D"].data_ptr(), N, L.EPI_BIAS
        a.bias = s["bias"].data_ptr()
        L.call("b2_gemm_bf16", a, stream.cuda_stream)


def adamw(stream, form):
    for b in range(NB):
        lo, hi = b * chunk, (b + 1) * chunk
        if form == "background":
            L.call("b2_adamw_background", g.data_ptr(), shadow.data_ptr(), master.data_ptr(), m.data_ptr(), v.data_ptr(),
                   decay.data_ptr(), lo, hi, hp, step_size.data_ptr(), stream.cuda_stream)
        else:
            L.call("b2_bucket_reduce_adamw", L.ptr_array([g.data_ptr()]), L.ptr_array([shadow.data_ptr()]), 1, 0,
                   master.data_ptr(), m.data_ptr(), v.data_ptr(), decay.data_ptr(), lo, hi, hp, step.data_ptr(),
                   stream.cuda_stream)


def run(do_gemm, form):
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        cur = torch.cuda.current_stream()
        s_gemm.wait_stream(cur); s_opt.wait_stream(cur)
        if do_gemm:
            gemms(s_gemm)
        if form:
            adamw(s_opt, form)
        cur.wait_stream(s_gemm); cur.wait_stream(s_opt)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


print("GEMM EW:", os.environ.get("B2_GEMM_EPI_WARPS", "16"))
tg = run(True, None)
for form in ("background", "regular"):
    ta = run(False, form)
    tb = run(True, form)
    print("%-10s gemm alone %.3f ms | adamw alone %.3f ms (%.2f TB/s) | together %.3f ms | hidden %.0f %% of the optimizer"
          % (form, tg, ta, n * 28 / ta / 1e9, tb, 100 * (tg + ta - tb) / ta), flush=True)
