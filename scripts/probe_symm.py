"""Probe (torchrun, >= 2 ranks): does torch symmetric memory work here, and what does a peer copy cost next to a running
attention kernel — copy engine (peer-buffer .copy_) vs NCCL all_gather?"""
import os, sys, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
import flashattn_b200 as fa
from flashattn_b200 import interface as I
I.set_shared_sms(True)
g = torch.Generator(device=dev).manual_seed(rank)
Q, K, V = (torch.randn(1, 8, 16384, 128, device=dev, generator=g).bfloat16() for _ in range(3))
def attn(): return fa.flash_attention_forward(Q, K, V, False)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(n):
        s, e = ev(), ev(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return round(sorted(ts)[len(ts) // 2], 3)
out["attn_alone_ms"] = timed(attn)
N = 128 << 20      # 256 MB of bf16
src = torch.randn(N, device=dev).bfloat16() if False else torch.empty(N, dtype=torch.bfloat16, device=dev).normal_()
# ---- NCCL all_gather next to the kernel
gathered = torch.empty(world * N, dtype=torch.bfloat16, device=dev)
def nccl_ag():
    w = dist.all_gather_into_tensor(gathered, src, async_op=True); w.wait()
out["nccl_all_gather_alone_ms"] = timed(nccl_ag)
def nccl_overlap():
    w = dist.all_gather_into_tensor(gathered, src, async_op=True); attn(); w.wait()
out["attn_plus_nccl_all_gather_ms"] = timed(nccl_overlap)
# ---- symmetric memory + copy engine
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(N, dtype=torch.bfloat16, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    peers = [hdl.get_buffer(r, (N,), torch.bfloat16) for r in range(world)]
    dst = torch.empty(world, N, dtype=torch.bfloat16, device=dev)
    side = [torch.cuda.Stream() for _ in range(min(world - 1, 4))]
    def ce_pull():
        cur = torch.cuda.current_stream(); e0 = ev(); e0.record(cur)
        for i in range(1, world):
            r = (rank + i) % world
            st = side[(i - 1) % len(side)]
            st.wait_event(e0)
            with torch.cuda.stream(st):
                dst[r].copy_(peers[r], non_blocking=True)
        for st in side:
            cur.wait_stream(st)
    out["ce_pull_alone_ms"] = timed(ce_pull)
    def ce_overlap():
        cur = torch.cuda.current_stream(); e0 = ev(); e0.record(cur)
        for i in range(1, world):
            r = (rank + i) % world
            st = side[(i - 1) % len(side)]
            st.wait_event(e0)
            with torch.cuda.stream(st):
                dst[r].copy_(peers[r], non_blocking=True)
        attn()
        for st in side:
            cur.wait_stream(st)
    out["attn_plus_ce_pull_ms"] = timed(ce_overlap)
    ok = all(torch.equal(dst[r][:1000], peers[r][:1000]) for r in range(world) if r != rank)
    out["ce_data_ok"] = bool(ok)
    out["gbps_ce_pull"] = round((world - 1) * N * 2 / out["ce_pull_alone_ms"] / 1e6, 1)
    out["gbps_nccl_ag"] = round((world - 1) * N * 2 / out["nccl_all_gather_alone_ms"] / 1e6, 1)
    t0 = time.perf_counter(); hdl.barrier(); torch.cuda.synchronize(); out["symm_barrier_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
except Exception as e:
    out["symm_error"] = repr(e)[:500]
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
