"""Data-parallel flow-matching step over NCCL: every rank runs the step (3x3 convs on the library) on its shard of the batch,
ONE all-reduce over the flattened gradients averages them (fm_train.average_param_grads); the result must equal the
single-GPU gradient of the full batch (same x_t, t, u; dropout off).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ddp_fm_check.py
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle.unet_module import UNet  # noqa: E402  (test infrastructure: the reference's module tree)
from ratio_guided_multimodal_fm_b200 import synthetic  # noqa: E402
from ratio_guided_multimodal_fm_b200.fm_train import library_convs, fm_mse, average_param_grads  # noqa: E402


def grads_of(model, x_t, t, u):
    for p in model.parameters():
        p.grad = None
    with library_convs(model):
        loss = fm_mse(model(x_t, t), u)
        loss.backward()
    return loss


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    for kind in ("unet_mnist32", "unet_svhn"):
        c = 3 if kind == "unet_svhn" else 1
        base = UNet(kind).eval()
        base.load_state_dict(synthetic.state_dict(kind, 12), strict=True)
        full, shard = copy.deepcopy(base).to(dev), copy.deepcopy(base).to(dev)
        per = 8
        B = per * world
        g = torch.Generator().manual_seed(13)
        x_t, t, u = torch.randn(B, c, 32, 32, generator=g).to(dev), torch.rand(B, generator=g).to(dev), torch.randn(B, c, 32, 32, generator=g).to(dev)
        grads_of(full, x_t, t, u)
        ref = torch.cat([p.grad.reshape(-1) for p in full.parameters()])
        lo = rank * per
        grads_of(shard, x_t[lo:lo + per], t[lo:lo + per], u[lo:lo + per])
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        average_param_grads(list(shard.parameters()))
        e1.record()
        torch.cuda.synchronize()
        got = torch.cat([p.grad.reshape(-1) for p in shard.parameters()])
        rel = ((got - ref).norm() / ref.norm()).item()
        tt = torch.tensor([rel], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{kind}: world={world} flat gradient {ref.numel() * 4 / 1e6:.1f} MB, flatten + all-reduce + scatter {e0.elapsed_time(e1):.3f} ms, "
                  f"relative L2 (averaged shard gradients vs full-batch gradient) {tt.item():.2e}")
        assert tt.item() <= 1e-4
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
