"""Raw host<->device copy ceiling of the box, no kernels and nothing of this repo in the path:
pinned cudaMemcpyAsync H2D and D2H through libcudart alone, one process per GPU under torchrun (torch.distributed is used
for the barrier and the max over ranks only).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_ceiling.py

Per rank: three 33 MB H2D buffers pairs + one 33 MB D2H in flight per lane, 3 lanes (the shape of bench.py's e2e leg for
c2_blend: 2 layers up, 1 result down per frame), plus the single-direction figures.  Variants: plain pinned memory,
write-combined upload staging, NUMA binding of the process to the GPU's node before allocating.
Prints one JSON line per (N, variant): aggregate GB/s over all ranks."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    numa = os.environ.get("PCIE_NUMA", "1") == "1"
    wc = os.environ.get("PCIE_WC", "0") == "1"
    cpus = None
    if numa:
        from zosimos_b200.shard import bind_to_gpu_numa
        cpus = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + ["libcudart.so.12"]
    rt = C.CDLL(cands[0])
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
    rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
    rt.cudaSetDevice(local)
    FB = 3840 * 2160 * 4
    LANES = 3

    def chk(e):
        if e != 0:
            raise RuntimeError("cuda error %d" % e)

    def host(n, flags):
        p = C.c_void_p()
        chk(rt.cudaHostAlloc(C.byref(p), n, flags))
        C.memset(p, 1, n)
        return p

    def dev(n):
        p = C.c_void_p()
        chk(rt.cudaMalloc(C.byref(p), n))
        return p
    lanes = []
    for _ in range(LANES):
        s = C.c_void_p(); chk(rt.cudaStreamCreate(C.byref(s)))
        lanes.append((s, host(2 * FB, 4 if wc else 0), host(FB, 0), dev(2 * FB), dev(FB)))  # 4 = cudaHostAllocWriteCombined

    def sync():
        for l in lanes:
            chk(rt.cudaStreamSynchronize(l[0]))
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(mode, n):
        sync()
        t0 = time.perf_counter()
        for i in range(n):
            s, hin, hout, din, dout = lanes[i % LANES]
            if mode in ("h2d", "both"):
                chk(rt.cudaMemcpyAsync(din, hin, 2 * FB, 1, s))
            if mode in ("d2h", "both"):
                chk(rt.cudaMemcpyAsync(hout, dout, FB, 2, s))
        for l in lanes:
            chk(rt.cudaStreamSynchronize(l[0]))
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt
    out = {"n_gpus": world, "numa_bound": bool(cpus), "write_combined_upload": wc, "frame_bytes": FB}
    for mode, per in (("h2d", 2 * FB), ("d2h", FB), ("both", 3 * FB)):
        run(mode, 12)
        n = 240
        dt = run(mode, n)
        out[mode + "_gbs_aggregate"] = round(per * n * world / dt / 1e9, 1)
        out[mode + "_gbs_per_gpu"] = round(per * n / dt / 1e9, 1)
    out["c2_blend_e2e_ceiling_mps"] = round(out["both_gbs_aggregate"] * 1e9 / 12 / 1e6, 0)  # 12 B cross PCIe per output pixel
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def single_process(ndev):
    """The same traffic from ONE process that owns a context on each of `ndev` devices ('one process with 8 contexts
    vs 8 processes'): copies are issued round robin over the devices' lanes from one host thread."""
    import glob
    import torch
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + ["libcudart.so.12"]
    rt = C.CDLL(cands[0])
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
    rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
    FB, LANES = 3840 * 2160 * 4, 3

    def chk(e):
        if e != 0:
            raise RuntimeError("cuda error %d" % e)
    lanes = []
    for dev in range(ndev):
        chk(rt.cudaSetDevice(dev))
        for _ in range(LANES):
            s, hin, hout, din, dout = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            chk(rt.cudaStreamCreate(C.byref(s)))
            chk(rt.cudaHostAlloc(C.byref(hin), 2 * FB, 1)); chk(rt.cudaHostAlloc(C.byref(hout), FB, 1))  # 1 = portable: pinned for every context
            C.memset(hin, 1, 2 * FB); C.memset(hout, 1, FB)
            chk(rt.cudaMalloc(C.byref(din), 2 * FB)); chk(rt.cudaMalloc(C.byref(dout), FB))
            lanes.append((dev, s, hin, hout, din, dout))

    def run(mode, n):
        for l in lanes:
            chk(rt.cudaSetDevice(l[0])); chk(rt.cudaStreamSynchronize(l[1]))
        t0 = time.perf_counter()
        for i in range(n):
            for k in range(ndev):  # one frame on every device per round
                dev, s, hin, hout, din, dout = lanes[k * LANES + i % LANES]
                chk(rt.cudaSetDevice(dev))
                if mode in ("h2d", "both"):
                    chk(rt.cudaMemcpyAsync(din, hin, 2 * FB, 1, s))
                if mode in ("d2h", "both"):
                    chk(rt.cudaMemcpyAsync(hout, dout, FB, 2, s))
        for l in lanes:
            chk(rt.cudaSetDevice(l[0])); chk(rt.cudaStreamSynchronize(l[1]))
        return time.perf_counter() - t0
    out = {"n_gpus": ndev, "single_process": True, "frame_bytes": FB}
    for mode, per in (("h2d", 2 * FB), ("d2h", FB), ("both", 3 * FB)):
        run(mode, 12)
        n = 240
        dt = run(mode, n)
        out[mode + "_gbs_aggregate"] = round(per * n * ndev / dt / 1e9, 1)
    out["c2_blend_e2e_ceiling_mps"] = round(out["both_gbs_aggregate"] * 1e9 / 12 / 1e6, 0)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--single-process":
        single_process(int(sys.argv[2]))
    else:
        main()
