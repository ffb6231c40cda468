"""2-GPU NCCL test of the data-parallel paths (eager hooks and CUDA-graph mode): after two optimizer steps on
different shards every rank must hold IDENTICAL parameters, and they must match a single-process run on the
concatenated batch.  Skipped on boxes with fewer than two GPUs."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TRANSFORMERS_OFFLINE="1")
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import hf_oracle as O
    from speechmix_b200 import SpeechMixEED, parallel
    from speechmix_b200.graph import GraphedTrainStep
    spc, txc = O.speech_config("mini"), O.text_config("bart-mini")
    if mode == "layerdrop":      # every rank draws its own LayerDrop decisions: gradient sets differ between ranks
        spc.layerdrop = 0.5

    def build():
        m = SpeechMixEED(spc, txc, down_scale=2)
        parallel.init_like_reference(m, seed=0)
        return m.to(dev).train()

    x, y = O.synthetic_batch(4, 1.0, 8, txc.vocab_size, seed=0)
    sl = parallel.shard_batch(4, rank, world)
    xs, ys = x[sl].to(dev), y[sl].to(dev)
    m = build()
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    red = parallel.GradientAllReducer(m, world, bucket_mb=1, payload="bf16" if mode == "bf16" else "fp32")
    if mode == "layerdrop":
        torch.manual_seed(100 + rank)
    if mode == "graph":
        g = GraphedTrainStep(m, opt, xs, ys, warmup=2, reducer=red)   # 2 eager steps + 1 replay = 3 steps
        g(xs, ys)
    else:
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            m(xs, labels=ys, return_model_detail=False)["loss"].backward()
            red.finish()
            opt.step()
    torch.cuda.synchronize()
    w = m.enc_to_dec_proj.weight.detach().clone()
    others = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(others, w)
    same = all(torch.equal(o, others[0]) for o in others)
    ok_ref = True
    if rank == 0 and mode != "layerdrop":
        ref = build()
        ropt = torch.optim.SGD(ref.parameters(), lr=0.05)
        xf, yf = x.to(dev), y.to(dev)
        for _ in range(3):
            ropt.zero_grad(set_to_none=True)
            ref(xf, labels=yf, return_model_detail=False)["loss"].backward()
            ropt.step()
        w0 = build().enc_to_dec_proj.weight.detach()
        d_ref, d_got = ref.enc_to_dec_proj.weight.detach() - w0, w - w0
        cos = float((d_ref * d_got).sum() / (d_ref.norm() * d_got.norm() + 1e-20))
        ok_ref = cos > (0.95 if mode == "bf16" else 0.98) and abs(float(d_got.norm() / d_ref.norm()) - 1.0) < 0.1
        ret["cos"] = cos
    ret[rank] = bool(same and ok_ref)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["eager", "graph", "bf16", "layerdrop"])
def test_two_gpu_data_parallel_matches_full_batch(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29640 + ["eager", "graph", "bf16", "layerdrop"].index(mode), mode, ret), nprocs=2, join=True)
    assert ret[0] and ret[1], dict(ret)
