"""Probe of the CUDA-IPC peer window under torchrun (2+ ranks): set-up time, copy-engine bandwidth into a peer's
buffer (cudaMemcpyAsync through the C ABI) and the push -> flag -> wait-kernel latency."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as td
import verde_b200 as vb
from verde_b200 import _cabi, dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
lib = _cabi.require_device(); lib.vb200_set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
td.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
t0 = time.perf_counter()
shared = {"buf": dist.SharedBuffer(64 * 1024 * 1024, torch.float64), "flags": dist.SharedBuffer(64, torch.int64)}
buf, flags = shared["buf"].tensor, shared["flags"].tensor
buf.zero_(); flags.zero_()
src_flag = torch.zeros(1, dtype=torch.int64, device=dev)
win = dist.PeerWindow(shared)
torch.cuda.synchronize(); td.barrier()
out = {"rank": rank, "world": world, "setup_s": round(time.perf_counter() - t0, 3)}
peer = (rank + 1) % world
comm = torch.cuda.Stream(priority=-1)
h = ctypes.c_void_p(comm.cuda_stream)
for nbytes in (8, 1 << 20, 100 << 20, 400 << 20):
    times = []
    for rep in range(4):
        td.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(comm)
        lib.vb200_peer_copy_async(ctypes.c_void_p(win.ptr(peer, "buf")), ctypes.c_void_p(buf.data_ptr()), nbytes, h)
        b.record(comm)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    out["copy_%d" % nbytes] = {"device_ms": round(min(times), 4), "GBps": round(nbytes / (min(times) * 1e-3) * 1e-9, 1)}
# all-to-one-neighbour pushes at once are above; now one sender to ALL peers in sequence (what a panel owner does)
if world > 2:
    times = []
    for rep in range(3):
        td.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(comm)
        if rank == 0:
            for r in range(1, world):
                lib.vb200_peer_copy_async(ctypes.c_void_p(win.ptr(r, "buf")), ctypes.c_void_p(buf.data_ptr()), 400 << 20, h)
        b.record(comm)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    out["one_to_all_400MB_ms"] = round(min(times), 3)
lat = []
for rep in range(5):
    epoch = 1000 + rep
    src_flag.fill_(epoch); torch.cuda.synchronize(); td.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if rank == 0:
        lib.vb200_peer_copy_async(ctypes.c_void_p(win.ptr(1, "buf")), ctypes.c_void_p(buf.data_ptr()), 1 << 20, h)
        lib.vb200_peer_copy_async(ctypes.c_void_p(win.ptr(1, "flags") + 24), ctypes.c_void_p(src_flag.data_ptr()), 8, h)
    elif rank == 1:
        lib.vb200_chol_wait_panel(ctypes.c_void_p(flags.data_ptr() + 24), epoch)
    b.record()
    torch.cuda.synchronize()
    lat.append(a.elapsed_time(b))
out["push_1MB_flag_wait_ms"] = [round(x, 3) for x in lat]
print(json.dumps(out), flush=True)
td.barrier()
td.destroy_process_group()
