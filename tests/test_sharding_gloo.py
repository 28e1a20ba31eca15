"""N > 1 host logic on CPU: world_size-2 gloo run of utterance-sharded inference (no data-path collective),
with the oracle standing in for the module (both backends share the API).  Each rank decodes its shard with
global utterance ids; the gathered result must equal the unsharded run."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import synthetic
    from transformer_tacotron2_b200.sharding import sharded_inference
    model = synthetic.make_model(stop_bias=-0.45)
    ph, pl, _, _ = synthetic.make_inputs(5, 14, 8, 91, ragged=True)
    lo, hi, out = sharded_inference(model, ph, pl, max_len=24, seed=7, rank=rank, world_size=world)
    # timing contract of bench.py: barrier on both sides, max over ranks
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, [x.numpy() for x in out]))      # host-side gather of results only
    if rank == 0:
        ma, lens, st = model.inference(ph, pl, max_len=24, seed=7)
        ok = True
        for lo_, hi_, (a, l, s) in gathered:
            T = a.shape[1]
            ok &= lens[lo_:hi_].tolist() == l.tolist()
            ok &= bool(torch.allclose(torch.from_numpy(a), ma[lo_:hi_, :T], atol=2e-5))
            ok &= bool((ma[lo_:hi_, T:] == 0).all())
            ok &= bool(torch.allclose(torch.from_numpy(s), st[lo_:hi_, :T], atol=2e-5))
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_inference_world2_gloo(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _dp_worker(rank, world, port, out_path):
    """Data-parallel training exchange: each rank differentiates its own shard (per-rank BatchNorm
    statistics, dropout masks keyed by the GLOBAL utterance id), ONE all-reduce over the flat gradient, mean = sum / world."""
    import copy
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import synthetic
    from oracle.transformer_tts import tts_loss
    from transformer_tacotron2_b200.sharding import shard_range
    from transformer_tacotron2_b200.training import allreduce_sum_
    base = synthetic.make_model(stop_bias=-8.0)
    ph, pl, mels, ml = synthetic.make_inputs(4, 6, 8, 17, ragged=True)

    def shard_grads(r):
        lo, hi = shard_range(4, world, r)
        m = copy.deepcopy(base).train()
        out = m(ph[lo:hi], pl[lo:hi], mels[lo:hi], ml[lo:hi], seed=5, utt_ids=list(range(lo, hi)))
        tts_loss(*out, mels[lo:hi], ml[lo:hi]).backward()
        return torch.cat([p.grad.flatten() for p in m.parameters()])

    flat = shard_grads(rank)
    allreduce_sum_(flat)
    flat /= world
    if rank == 0:
        want = sum(shard_grads(r) for r in range(world)) / world
        ok = bool(torch.allclose(flat, want, atol=1e-6, rtol=1e-5))
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_world2_gloo(tmp_path):
    out = str(tmp_path / "dp.txt")
    port = 30100 + (os.getpid() % 500)
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_ranges_cover_and_balance():
    from transformer_tacotron2_b200.sharding import shard_ranges
    for n in (1, 5, 64, 67):
        for w in (1, 2, 4, 8):
            r = shard_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
