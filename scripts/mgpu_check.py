"""Real multi-GPU parity check of the strip-sharded path (run under torchrun, one rank per GPU):
every rank filters its strip through stage1 / NCCL all_gather / stage2; rank 0 also filters the whole
image unsharded and all strips are compared with it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import recfilter_b200 as rf
from recfilter_b200 import Plan, Scan, gaussian_weights
from recfilter_b200.sharded import ShardedFilter

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); rf.lib().rf_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
G3 = gaussian_weights(5.0, 3)
worst = 0.0
for (W, H) in [(2048, 2048), (8192, 8192)]:
    scans = [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)]
    gen = torch.Generator(device="cuda").manual_seed(99)
    full = torch.rand((H, W), device="cuda", generator=gen)            # same image on every rank
    flt = ShardedFilter((W, H), "f32", scans, "clamp", rank=rank, world=world, shard_dim=1, batch=1,
                        exchange=os.environ.get("RF_EXCHANGE", "auto"))
    lo, hi = flt.rows
    src = full[lo:hi].contiguous(); dst = torch.empty_like(src)
    flt.run([src], [dst]); torch.cuda.synchronize()
    ref_plan = Plan((W, H), "f32", scans, "clamp")
    ref = torch.empty_like(full); ref_plan.execute(full.view(-1), ref.view(-1)); torch.cuda.synchronize()
    err = (dst - ref[lo:hi]).abs().max() / ref.abs().max()
    t = torch.tensor([float(err)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = max(worst, float(t.item()))
    if rank == 0: print(f"sharded vs unsharded {W}x{H} over {world} GPUs: max rel diff {float(t.item()):.3e}", flush=True)
    flt.close(); ref_plan.close()
# stacked batch: B images as one [B][rows][W] stack, one launch sequence and one exchange
B, W, H = 3, 2048, 2048
scans = [Scan(0, True, G3), Scan(0, False, G3), Scan(1, True, G3), Scan(1, False, G3)]
gen = torch.Generator(device="cuda").manual_seed(7)
full = torch.rand((B, H, W), device="cuda", generator=gen)
flt = ShardedFilter((W, H), "f32", scans, "clamp", rank=rank, world=world, shard_dim=1, batch=B, stacked=True,
                    exchange=os.environ.get("RF_EXCHANGE", "auto"))
if rank == 0: print("exchange:", "p2p windows, neighbours only" if flt.neighbor else "p2p windows" if flt.p2p else ("alltoall (column-chunked)" if flt.chunked else "allgather"), flush=True)
lo, hi = flt.rows
src = full[:, lo:hi].contiguous(); dst = torch.empty_like(src)
flt.run_stacked(src, dst); torch.cuda.synchronize()
ref_plan = Plan((W, H), "f32", scans, "clamp")
for b in range(B):
    ref = torch.empty_like(full[b]); ref_plan.execute(full[b].contiguous().view(-1), ref.view(-1)); torch.cuda.synchronize()
    err = (dst[b] - ref[lo:hi]).abs().max() / ref.abs().max()
    t = torch.tensor([float(err)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = max(worst, float(t.item()))
if rank == 0: print(f"stacked batch of {B} over {world} GPUs: worst rel diff so far {worst:.3e}", flush=True)
flt.close(); ref_plan.close()
assert worst <= 2e-5, worst
if rank == 0: print("MGPU PARITY OK", flush=True)
dist.destroy_process_group()
