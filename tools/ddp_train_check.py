"""Data-parallel ratio-estimator training over NCCL: every rank runs rgfm_ratio_trainer_step on its shard, ONE all-reduce
over the flat gradient arena averages the gradients; the result must equal the single-GPU gradient of the full batch
(disc loss, equal real / fake counts per shard, no dropout).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_train_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from ratio_guided_multimodal_fm_b200 import RatioTrainPlan, synthetic  # noqa: E402
from ratio_guided_multimodal_fm_b200.train import average_gradients  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for kind in ("ratio28", "ratio_mnist_svhn"):
        sd = synthetic.state_dict(kind, 21)
        params = {k: v.to(dev).contiguous() for k, v in sd.items() if v.dtype.is_floating_point}
        per = 8
        B = per * world
        g = torch.Generator().manual_seed(11)
        hw, cy = (28, 1) if kind == "ratio28" else (32, 3)
        x, y = torch.randn(B, 1, hw, hw, generator=g).to(dev), torch.randn(B, cy, hw, hw, generator=g).to(dev)
        r = (torch.arange(B) % 2 == 0).float().to(dev)
        full = RatioTrainPlan(kind, params, max_batch=B, device=dev)
        _, gfull = full.step(x, y, r, "disc")
        ref = full.flat_grad.clone()
        shard = RatioTrainPlan(kind, params, max_batch=per, device=dev)
        lo = rank * per
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        shard.step(x[lo:lo + per], y[lo:lo + per], r[lo:lo + per], "disc")
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        average_gradients(shard.flat_grad)
        e1.record()
        torch.cuda.synchronize()
        rel = ((shard.flat_grad - ref).norm() / ref.norm()).item()
        t = torch.tensor([rel], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{kind}: world={world} flat gradient {ref.numel() * 4 / 1e6:.1f} MB, all-reduce {e0.elapsed_time(e1):.3f} ms, "
                  f"relative L2 (averaged shard gradients vs full-batch gradient) {t.item():.2e}")
        assert t.item() <= 1e-4
        full.close(); shard.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
