"""Multi-GPU substrate probe (torchrun): symmetric heap modes, P2P read bandwidth from inside a
kernel, NVLS multicast store, flag latency.  Prints one JSON line from rank 0."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json, os, sys, time
import torch, torch.distributed as dist
from bflc_demo_b200._native import C
from bflc_demo_b200.parallel.symm import SymmetricHeap

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    m = C()
    out = {"world": world, "mc_supported": bool(m.SymmHeap.multicast_supported(lr))}
    nbytes = 512 << 20
    for mode in ("vmm", "ipc"):
        try:
            h = SymmetricHeap(nbytes, rank=rank, world=world, device=lr, mode=mode)
        except Exception as e:  # noqa
            out[mode] = {"error": repr(e)[:300]}
            continue
        info = h.describe()
        n_vec = (256 << 20) // 16
        mine = h.view(0, [n_vec * 4], torch.float32)
        mine.fill_(float(rank + 1))
        dst = torch.empty(n_vec * 4, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        peer = (rank + 1) % world
        src_ptr = h.peer_ptrs[peer]
        for _ in range(2): m.p2p_read_probe(src_ptr, dst.data_ptr(), n_vec)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m.p2p_read_probe(src_ptr, dst.data_ptr(), n_vec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        info["p2p_read_GBps"] = (256 << 20) / ms / 1e6
        info["p2p_read_ok"] = bool((dst[:1024] == float(peer + 1)).all().item())
        # local copy for reference
        loc = h.view(256 << 20, [n_vec * 4], torch.float32)
        e0.record()
        for _ in range(5): m.p2p_read_probe(h.local_ptr, loc.data_ptr(), n_vec)
        e1.record(); torch.cuda.synchronize()
        info["local_read_GBps"] = (256 << 20) / (e0.elapsed_time(e1) / 5) / 1e6
        if h.has_multicast:
            dist.barrier()
            srcb = torch.full((n_vec * 4,), 7.0 + rank, device="cuda")
            if rank == 0:
                m.mc_store_probe(h.mc_ptr, srcb.data_ptr(), n_vec)
            torch.cuda.synchronize(); dist.barrier()
            info["mc_store_landed_everywhere"] = bool((mine[:4096] == 7.0).all().item())
            dist.barrier()
            e0.record()
            if rank == 0:
                for _ in range(5): m.mc_store_probe(h.mc_ptr, srcb.data_ptr(), n_vec)
            e1.record(); torch.cuda.synchronize()
            info["mc_store_GBps_rank0"] = (256 << 20) / (e0.elapsed_time(e1) / 5) / 1e6
        allinfo = [None] * world
        dist.all_gather_object(allinfo, info)
        out[mode] = allinfo[0]
        out[mode]["p2p_read_GBps_min_over_ranks"] = min(i["p2p_read_GBps"] for i in allinfo)
        out[mode]["all_ok"] = all(i["p2p_read_ok"] for i in allinfo)
        del mine, loc, h
        torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(out))
    dist.barrier(); dist.destroy_process_group()

main()
