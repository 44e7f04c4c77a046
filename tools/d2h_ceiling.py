#!/usr/bin/env python
"""Bare pinned-memory copy ceiling of the box: N processes (one per GPU) each stream a large device buffer into pinned host
memory (D2H) and back (H2D) with plain cudaMemcpyAsync, nothing else running.  This is the ceiling the e2e decode path
(47 GB of PCM out per GPU per step) is measured against.

    python tools/d2h_ceiling.py --procs 1,2,4,8 [--gb 4] [--seconds 3] [--streams 1,2]

Prints one JSON line per (N, copy streams) with per-GPU and aggregate GB/s, the NUMA node of every GPU and the CPU affinity.
"""
import argparse
import json
import os
import subprocess
import sys
import time


def numa_of_gpu(i):
    try:
        import torch
        bus = torch.cuda.get_device_properties(i).pci_bus_id
        dom = torch.cuda.get_device_properties(i).pci_domain_id
        dev = torch.cuda.get_device_properties(i).pci_device_id
        p = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        return int(open(p).read().strip())
    except Exception:
        return None


def worker(rank, world, gb, seconds, nstreams, bind, q_in, q_out):
    import torch
    torch.cuda.set_device(rank)
    node = numa_of_gpu(rank)
    if bind and node is not None and node >= 0:
        try:
            cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
            ids = []
            for part in cpus.split(","):
                a, _, b = part.partition("-")
                ids += list(range(int(a), int(b or a) + 1))
            os.sched_setaffinity(0, ids)
        except Exception:
            pass
    nbytes = int(gb * (1 << 30))
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.zero_()
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    part = nbytes // nstreams
    res = {}
    for direction in ("d2h", "h2d"):
        def one_pass():
            for k, s in enumerate(streams):
                with torch.cuda.stream(s):
                    a, b = k * part, (k + 1) * part
                    if direction == "d2h":
                        h[a:b].copy_(d[a:b], non_blocking=True)
                    else:
                        d[a:b].copy_(h[a:b], non_blocking=True)
        one_pass()
        torch.cuda.synchronize()
        q_out.put(("ready", rank))
        q_in.get()                      # start barrier
        t0 = time.perf_counter()
        moved = 0
        while time.perf_counter() - t0 < seconds:
            one_pass()
            torch.cuda.synchronize()
            moved += part * nstreams
        dt = time.perf_counter() - t0
        res[direction] = moved / dt / 1e9
    q_out.put(("done", rank, res, node, sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", default="1")
    ap.add_argument("--gb", type=float, default=4.0)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--streams", default="1,2")
    ap.add_argument("--bind", type=int, default=0, help="1: pin each process to the CPUs of its GPU's NUMA node")
    args = ap.parse_args()
    import torch
    import torch.multiprocessing as mp
    mp.set_start_method("spawn", force=True)
    ngpu = torch.cuda.device_count()
    for n in [int(x) for x in args.procs.split(",")]:
        if n > ngpu:
            print(json.dumps({"procs": n, "skipped": "only %d GPUs visible" % ngpu}), flush=True)
            continue
        for ns in [int(x) for x in args.streams.split(",")]:
            q_in, q_out = mp.Queue(), mp.Queue()
            ps = [mp.Process(target=worker, args=(r, n, args.gb, args.seconds, ns, args.bind, q_in, q_out)) for r in range(n)]
            for p in ps:
                p.start()
            out = {}
            for phase in range(2):
                for _ in range(n):
                    q_out.get()
                for _ in range(n):
                    q_in.put(1)
            for _ in range(n):
                m = q_out.get()
                while m[0] != "done":
                    m = q_out.get()
                out[m[1]] = m[2:]
            for p in ps:
                p.join()
            d2h = [out[r][0]["d2h"] for r in range(n)]
            h2d = [out[r][0]["h2d"] for r in range(n)]
            print(json.dumps({"procs": n, "copy_streams": ns, "bind_numa": args.bind, "gb_per_copy": args.gb,
                              "d2h_gbs_per_gpu": [round(x, 2) for x in d2h], "d2h_gbs_total": round(sum(d2h), 2),
                              "h2d_gbs_per_gpu": [round(x, 2) for x in h2d], "h2d_gbs_total": round(sum(h2d), 2),
                              "gpu_numa_nodes": [out[r][1] for r in range(n)], "cpus_per_proc": [out[r][3] for r in range(n)],
                              "host_cpus": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
