"""One rank per GPU: the peer-HBM DDP path vs the oracle's DDP restatement.  Exits non-zero on any mismatch.
Two launchers, as in the reference:
  * torchrun / torch.distributed.launch, env:// rendezvous (multi-gpu-distributed-cls.py:268-277, README.md:84):
      python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/ddp_worker.py
  * torch.multiprocessing.spawn with a tcp:// rendezvous (multi-gpu-distributed-mp-cls.py:265,361):
      python tests/ddp_worker.py --spawn 2
Three step loops: "eager" (fwd / criterion / zero_grad / backward / step, :166-176), "amp" (the -amp scripts' loop:
autocast + GradScaler.scale(loss).backward() / scaler.step / scaler.update, multi-gpu-distributed-mp-amp-cls.py
:166-171) and "fused" (the whole step as one CUDA graph).
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist
import torch.nn.functional as F

from parity import TOL_TRAJ, b2, bert_ref, state_from_hf_init, tiny_config
from oracle import ddp_ref


def main(local=None, world=None, init_method=None):
    if init_method is None:     # torchrun
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
    else:                       # mp.spawn(main_worker, nprocs=world, args=(world,)): first argument is the process index
        rank = local
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", init_method=init_method, world_size=world, rank=rank)
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    state = state_from_hf_init(cfg, seed=123)
    steps = 6
    batches = [[bert_ref.synthetic_batch(cfg, 4, 128, 7000 + 10 * s + r, padded=(s % 2 == 1)) for r in range(world)]
               for s in range(steps)]
    ref = {k: v.clone() for k, v in state.items()}
    hist = ddp_ref.train(ref, cfg, batches)

    class A:
        weight_decay, learning_rate = 0.01, 3e-5

    for mode in ("eager", "amp", "fused"):
        # every rank but 0 starts from different weights: the wrap-time broadcast must make rank 0 win
        init = state if rank == 0 else state_from_hf_init(cfg, seed=999)
        model = b2.BertForSequenceClassification(cfg)
        model.load_state_dict(init)
        model.cuda()
        ddp = b2.DistributedDataParallel(model, device_ids=[local])
        opt = b2.build_optimizer(ddp, A)
        assert [n for n, _ in ddp.named_parameters()][0].startswith("module.")
        fused = b2.FusedTrainStep(ddp, opt, 4, 128) if mode == "fused" else None
        scaler = torch.amp.GradScaler("cuda") if mode == "amp" else None
        for s in range(steps):
            b = batches[s][rank]
            if fused is not None:
                loss = fused(b)
            elif scaler is not None:
                d = {k: v.to(dev) for k, v in b.items()}
                with torch.autocast("cuda"):
                    out = ddp(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"],
                              attention_mask=d["attention_mask"], labels=d["label"])
                    loss = F.cross_entropy(out[1], d["label"])
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                assert float(scaler.get_scale()) == 65536.0   # bf16 gradients never trip the inf check
            else:
                d = {k: v.to(dev) for k, v in b.items()}
                out = ddp(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"],
                          attention_mask=d["attention_mask"], labels=d["label"])
                loss = F.cross_entropy(out[1], d["label"])
                opt.zero_grad()
                loss.backward()
                opt.step()
            red = ddp.loss_reduce(loss)
            lv, rv = float(loss), float(red)
            assert abs(lv - float(hist[s]["loss_per_rank"][rank])) <= TOL_TRAJ, (mode, s, rank, lv)
            assert abs(rv - float(hist[s]["loss_mean"])) <= TOL_TRAJ, (mode, s, rank, rv)
        if mode == "amp":
            # one rank's loss overflows: stock DDP all-reduces the gradients before GradScaler looks at them, so EVERY
            # rank skips the step and backs its scale off.  Here the inf check sees the classifier.bias probe, made
            # rank-consistent by a scalar peer exchange (ddp.consensus_probe): same outcome, no rank diverges.
            before = model._flat.clone()
            t_before = int(opt._state()["step"])
            d = {k: v.to(dev) for k, v in batches[0][rank].items()}
            with torch.autocast("cuda"):
                out = ddp(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"],
                          attention_mask=d["attention_mask"], labels=d["label"])
                loss = F.cross_entropy(out[1], d["label"]) * (float("inf") if rank == world - 1 else 1.0)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            torch.cuda.synchronize()
            assert float(scaler.get_scale()) == 32768.0, (rank, float(scaler.get_scale()))
            assert int(opt._state()["step"]) == t_before, rank
            assert torch.equal(model._flat, before), (rank, "a skipped step changed the weights")
            assert bool(torch.isfinite(model._engine.shadow.float()).all())
        # the reference's checkpoint idiom: ONLY rank 0 calls state_dict() (`if local_rank == 0: torch.save(
        # self.model.state_dict(), ...)`, multi-gpu-distributed-cls.py:190-197) -- it must not need the other ranks
        # (fp32 slices owned by peers are pulled one-sidedly out of their HBM)
        ckpt = os.path.join(tempfile.gettempdir(), "b2_ddp_worker_%s_%d.pt" % (mode, os.getppid()))
        if rank == 0:
            sd = ddp.state_dict()
            assert all(k.startswith("module.") for k in sd)
            for k, v in ref.items():
                err = float((sd["module." + k].cpu() - v).abs().max())
                assert err <= 2e-4, (mode, k, err)
            torch.save(sd, ckpt)
        torch.cuda.synchronize()
        dist.barrier()
        # ... and the reference's test phase (:357-363): fresh model -> cuda -> DDP wrap -> load_state_dict(torch.load(ckpt))
        # on the WRAPPED model -> eval.  The kernels must see the loaded weights (bf16 shadow refreshed), on every rank.
        fresh = b2.BertForSequenceClassification(cfg)
        fresh.load_state_dict(state_from_hf_init(cfg, seed=555))
        fresh.cuda()
        ddp2 = b2.DistributedDataParallel(fresh, device_ids=[local])
        eb = {k: v.to(dev) for k, v in batches[0][rank].items()}
        ddp2.eval()
        with torch.no_grad():
            before = ddp2(input_ids=eb["input_ids"], token_type_ids=eb["token_type_ids"],
                          attention_mask=eb["attention_mask"]).logits.clone()
        res = ddp2.load_state_dict(torch.load(ckpt))
        assert not res.missing_keys and not res.unexpected_keys, res
        with torch.no_grad():
            after = ddp2(input_ids=eb["input_ids"], token_type_ids=eb["token_type_ids"],
                         attention_mask=eb["attention_mask"]).logits.clone()
        _, want = bert_ref.forward(ref, cfg, batches[0][rank]["input_ids"], batches[0][rank]["token_type_ids"],
                                   batches[0][rank]["attention_mask"])
        assert float((after.cpu() - want).abs().max()) <= 1e-2, (mode, "loaded weights not in effect")
        assert float((before.cpu() - want).abs().max()) > 1e-2, (mode, "test is vacuous")
        sd2 = ddp2.state_dict()
        for k, v in ref.items():
            assert float((sd2["module." + k].cpu() - v).abs().max()) <= 2e-4, (mode, k)
        dist.barrier()
        if rank == 0:
            os.remove(ckpt)
        del ddp2, fresh
        # every rank holds identical bf16 weights after the exchange
        chk = model._engine.shadow.float().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(float(c) == float(allc[0]) for c in allc), allc
        # eval-time gather (Trainer.output_reduce)
        t = torch.full((4, 6), float(rank), device=dev)
        g = ddp.all_gather_rows(t)
        assert g.shape == (4 * world, 6)
        for r in range(world):
            assert float(g[4 * r:4 * r + 4].mean()) == float(r)
        lab = torch.arange(4, device=dev) + 100 * rank
        gl = ddp.all_gather_rows(lab)
        assert gl.tolist() == [i + 100 * r for r in range(world) for i in range(4)]
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print("ddp_worker: mode %s OK (world %d)" % (mode, world), flush=True)
        del fused, opt, ddp, model
    dist.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--spawn":
        import torch.multiprocessing as mp
        n = int(sys.argv[2])
        mp.spawn(main, nprocs=n, args=(n, "tcp://127.0.0.1:%d" % (29600 + os.getpid() % 300)))
    else:
        main()
