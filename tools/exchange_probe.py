"""Where the N-GPU gradient exchange spends its time, piece by piece (development aid for DESIGN.md §9 item 3).

Run under torchrun on one node, e.g.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
        tools/exchange_probe.py
For every bucket (embeddings | layer i | head) rank 0 prints the device time (CUDA events on the exchange stream, mean
of ITERS repetitions after warm-up, max over ranks) of: the flag barrier, the copy-engine pull of the peers' slices,
the local reduce + AdamW kernel, the copy-engine push of the new bf16 slice, and -- for comparison -- the kernel form
(`b2_bucket_reduce_adamw` reading and writing the peers through mapped pointers).  Nothing else runs on the GPUs, so
these are lower bounds of what the pieces cost inside a step (where they time-slice with the GEMM CTAs).
Written at the end of round 1 without a GPU at hand: not yet exercised.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pytorch_distributed_nlp_b200 as b2  # noqa: E402
from pytorch_distributed_nlp_b200 import _lib as L  # noqa: E402
from pytorch_distributed_nlp_b200 import ddp as ddp_mod  # noqa: E402

ITERS, WARM = 20, 5


def timed(stream, fn):
    """mean device milliseconds of fn() on `stream`"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(WARM):
        fn()
    stream.synchronize()
    e0.record(stream)
    for _ in range(ITERS):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / ITERS


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = b2.chinese_bert_wwm_ext_config(num_labels=6)
    b2.set_seed(123)
    model = b2.BertForSequenceClassification(cfg).cuda()
    wrap = b2.DistributedDataParallel(model, device_ids=[local])

    class A:
        weight_decay, learning_rate = 0.01, 3e-5

    opt = b2.build_optimizer(wrap, A)
    comm = wrap.comm
    side = wrap._side
    s = side.cuda_stream
    peers_g, peers_s = comm.peers["grads"], comm.peers["shadow"]
    rows = []
    with torch.cuda.stream(side):
        for idx, ((b0, e0, label), (sb, se)) in enumerate(zip(model._layout.buckets, wrap._slices)):
            nbytes = 2 * (se - sb)
            slot = ddp_mod._SLOT_BUCKET0 + idx

            def barrier():
                comm.barrier(slot, s)

            def pull():
                for r in range(world):
                    if r != rank:
                        L.call("b2_copy_async", wrap._stage.data_ptr() + wrap._stage_off[idx][r], peers_g[r] + 2 * sb,
                               nbytes, s)

            g_local, s_local = [], []
            for r in range(world):
                if r == rank:
                    g_local.append(peers_g[r])
                    s_local.append(peers_s[r])
                else:
                    g_local.append(wrap._stage.data_ptr() + wrap._stage_off[idx][r] - 2 * sb)
                    s_local.append(None)

            def reduce_local():
                opt.update_range(sb, se, world, rank, g_local, s_local, s)

            def push():
                for r in range(world):
                    if r != rank:
                        L.call("b2_copy_async", peers_s[r] + 2 * sb, peers_s[rank] + 2 * sb, nbytes, s)

            def kernel_form():
                opt.update_range(sb, se, world, rank, peers_g, peers_s, s)

            t = []
            for fn in (barrier, pull, reduce_local, push, kernel_form):
                dist.barrier()
                t.append(timed(side, fn))
            tt = torch.tensor(t, dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            rows.append((label, se - sb, [float(x) * 1e3 for x in tt]))
    if rank == 0:
        print("world %d, per-bucket device time in us (max over ranks): slice elements | barrier | DMA pull | "
              "local reduce+AdamW | DMA push | kernel form" % world)
        tot = [0.0] * 5
        for label, n, t in rows:
            print("%-12s %10d | %7.1f | %7.1f | %7.1f | %7.1f | %7.1f" % ((label, n) + tuple(t)))
            tot = [a + b for a, b in zip(tot, t)]
        print("%-12s %10s | %7.1f | %7.1f | %7.1f | %7.1f | %7.1f" % (("TOTAL", "") + tuple(tot)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
