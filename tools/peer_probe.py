"""tools/peer_probe.py -- NVLink push bandwidth of an SM kernel, variants of tools/peer_probe.cu.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py

Every rank pushes n rows of 64 bytes into its right neighbour's symmetric buffer (all ranks at once, like the
sharded step), CUDA-event timed; also a local run (dst = own memory) and cudaMemcpyPeer for reference."""
import ctypes
import os
import subprocess
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpeer_probe.so")


def build():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "peer_probe.cu")):
        subprocess.run(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                        "-o", SO, os.path.join(HERE, "peer_probe.cu")], check=True)
    return SO


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = ctypes.CDLL(SO)
    lib.probe_launch.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                 ctypes.c_int, ctypes.c_void_p]
    n = 1_500_000                                          # rows of 64 B: 96 MB, the cfg2 payload of one rank
    n_src = 5_000_000
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(n * 16, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        peer = hdl.buffer_ptrs[(rank + 1) % world]
    else:
        buf = torch.empty(n * 16, dtype=torch.float32, device=dev)
        peer = buf.data_ptr()
    src = torch.randn(n_src * 32, device=dev)
    ids = torch.randint(0, n_src, (n,), device=dev, dtype=torch.int32)
    own = torch.empty(n * 16, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, dst in (("local", own.data_ptr()), ("peer", peer)):
        for variant in range(4):
            for ctas in (148 * 2, 148 * 8):
                for _ in range(3):
                    lib.probe_launch(variant, src.data_ptr(), ids.data_ptr(), dst, n, ctas, st)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    lib.probe_launch(variant, src.data_ptr(), ids.data_ptr(), dst, n, ctas, st)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 100.0
                out["%s v%d ctas%d" % (name, variant, ctas)] = (round(us, 1), round(n * 64 / us * 1e-3, 1))
    # correctness of the bulk gather variant against the LSU one (local)
    lib.probe_launch(1, src.data_ptr(), ids.data_ptr(), own.data_ptr(), n, 296, st)
    a = own.clone()
    own.zero_()
    lib.probe_launch(3, src.data_ptr(), ids.data_ptr(), own.data_ptr(), n, 296, st)
    torch.cuda.synchronize()
    out["bulk_gather_equals_lsu_gather"] = bool(torch.equal(a, own))
    # PULL: random 64-byte rows read straight from the PEER's table (variant 1 with src = peer memory, dst = own)
    if world > 1:
        tbl = symm_mem.empty(n_src * 32, dtype=torch.float32, device=dev)
        th = symm_mem.rendezvous(tbl, dist.group.WORLD)
        tbl.copy_(src)
        torch.cuda.synchronize()
        dist.barrier()
        peer_tbl = th.buffer_ptrs[(rank + 1) % world]
        for name, sp in (("pull local", tbl.data_ptr()), ("pull peer", peer_tbl)):
            for variant in (1, 3, 0):
                for ctas in (148 * 4, 148 * 8, 148 * 16):
                    for _ in range(3):
                        lib.probe_launch(variant, sp, ids.data_ptr(), own.data_ptr(), n, ctas, st)
                    torch.cuda.synchronize()
                    dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        lib.probe_launch(variant, sp, ids.data_ptr(), own.data_ptr(), n, ctas, st)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 100.0
                    out["%s v%d ctas%d" % (name, variant, ctas)] = (round(us, 1), round(n * 64 / us * 1e-3, 1))
    if world > 1:
        dst_t = hdl.get_buffer((rank + 1) % world, (n * 16,), torch.float32)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst_t.copy_(own)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        out["memcpy_peer"] = (round(us, 1), round(n * 64 / us * 1e-3, 1))
    if rank == 0:
        for k, v in out.items():
            print(k, v, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    build()
    main()
