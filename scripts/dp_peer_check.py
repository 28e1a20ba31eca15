"""2+ GPU check of the fused data-parallel optimiser step (torchrun --nproc-per-node N scripts/dp_peer_check.py):
the peer-memory kernel (reduce-scatter -> Adam -> all-gather over NVLink) must leave every rank with the parameters the
NCCL all-reduce + Adam path produces, and the two paths are timed against each other."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from transformer_tacotron2_b200 import TransformerTTS  # noqa: E402
from transformer_tacotron2_b200.training import Trainer  # noqa: E402


def main():
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    dist.init_process_group("nccl", device_id=dev)
    src = bench.synthetic_state_dict()
    B, S, T = 8, 40, 120
    g = torch.Generator().manual_seed(100 + rank)
    ph = torch.randint(1, 128, (B, S), generator=g).to(dev); pl = torch.full((B,), S, dtype=torch.int32, device=dev)
    mel = torch.randn(B, T, 80, generator=g).clamp(-4, 4).to(dev); ml = torch.full((B,), T, dtype=torch.int32, device=dev)
    trs = {}
    for fused in (False, True):
        m = TransformerTTS(device=lr_); m.load_state_dict(src.state_dict())
        trs[fused] = Trainer(m, lr=1e-3, world_size=world, rank=rank, fused_peer_adam=fused)
        assert trs[fused]._peers == fused, "peer mapping failed"
        print(f"[rank {rank}] trainer fused={fused} ready", flush=True)
    for step in range(3):
        # identical gradients for both paths (two separate backward passes differ in the last bits -- atomic accumulation order --
        # and early Adam steps are sign-like, which would turn those bits into +-lr differences)
        trs[False].forward_backward(ph, pl, mel, ml, seed=5 + step, utt_offset=rank * B)
        trs[True].flat_grads.copy_(trs[False].flat_grads)
        trs[False].all_reduce_grads(); trs[False].adam_step()
        torch.cuda.synchronize(); print(f"[rank {rank}] step {step}: nccl path done", flush=True)
        trs[True].adam_step_peers()
        torch.cuda.synchronize(); print(f"[rank {rank}] step {step}: peer path done", flush=True)
        # keep the two parameter sets identical going into the next step's gradient computation
    out = {}
    for fused in (False, True):
        flat = torch.cat([v.flatten() for _, v in sorted(trs[fused].parameters().items())])
        out[fused] = flat
        ref = flat.to(dev).clone(); dist.broadcast(ref, 0)
        assert torch.equal(ref.cpu(), flat), f"rank {rank}: parameters differ from rank 0 (fused={fused})"
    err = float((out[True] - out[False]).abs().max())
    rel = float((out[True] - out[False]).norm() / out[False].norm())
    if rank == 0:
        print(f"after 3 steps on identical gradients: max |fused - nccl| = {err:.3e}, rel-L2 = {rel:.3e}")
    assert rel < 1e-6, rel
    for fused in (False, True):                      # timing of the optimiser part only
        tr = trs[fused]
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            if fused:
                tr.adam_step_peers()
            else:
                tr.all_reduce_grads(); tr.adam_step()
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"fused={fused}: exchange + Adam + repack {e0.elapsed_time(e1) / 20:.3f} ms per step (world {world})")
    dist.barrier(); dist.destroy_process_group()
    if rank == 0:
        print("dp_peer_check ok")


if __name__ == "__main__":
    main()
