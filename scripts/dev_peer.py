"""torchrun, 2+ ranks: gather variant over PeerCollectives at growing sizes, synchronising and printing after every phase."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
import flashattn_b200 as fa
import flashattn_b200.sharding as sh
from flashattn_b200 import _cabi
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def log(*a):
    print(f"[r{rank} {time.strftime('%H:%M:%S')}]", *a, file=sys.stderr, flush=True)
coll = sh.PeerCollectives()
log("coll ok")
for (H, S2, groups) in ((4, 2048, [1, 3]), (32, 4096, None), (32, 16384, None), (32, 65536, None), (32, 131072 // world, None)):
    g = torch.Generator(device=dev).manual_seed(rank)
    q, k, v, do = (torch.randn(1, H, S2, 128, device=dev, generator=g).bfloat16() for _ in range(4))
    for step in range(3):
        O, LSE, saved = sh.gather_attention_forward(q, k, v, None, None, coll, groups)
        torch.cuda.synchronize(); log(H, S2, "step", step, "fwd ok", _cabi.last_hang())
        dq, dk, dv = sh.gather_attention_backward(q, O, do, LSE, saved, None, None, coll)
        torch.cuda.synchronize(); log(H, S2, "step", step, "bwd ok", float(dk.float().abs().mean()), _cabi.last_hang())
    dist.barrier()
    # through autograd (backward on the autograd thread)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    for step in range(2):
        O = sh.gather_flash_attention(q, k, v, None, coll, groups); O.backward(do)
        torch.cuda.synchronize(); log(H, S2, "autograd step", step, "ok")
        q.grad = None; k.grad = None; v.grad = None
log("done")
dist.barrier(); dist.destroy_process_group()
