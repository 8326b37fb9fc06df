"""Host placement for the host-buffer path (``TokenPackerB200.forward_host`` / ``tp_forward_host``).

That path is bound by the host->device copies (377 MB in, 75 MB out per 64-crop call).  On a two-socket box the pinned
staging buffers of a rank must live on the NUMA node its GPU hangs off, or every copy crosses the socket interconnect and
eight ranks contend for it.  ``bind_to_gpu_node`` pins the calling process to the CPUs of that node; Linux's default
first-touch policy then places every later pinned allocation (``tensor.pin_memory()``, ``cudaHostAlloc``) there.  Call it once
per rank BEFORE allocating host buffers.  Pure sysfs + sched_setaffinity: no libnuma dependency, and a no-op (reported as such)
where the topology is not exposed (containers without /sys/bus/pci, single-node hosts).
"""
from __future__ import annotations

import os


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_address(device_index: int) -> str | None:
    """PCI address ``dddd:bb:dd.f`` of a CUDA device (through torch's device properties; nvidia-smi as a fallback)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        pass
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", f"--id={device_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        if out:
            dom, rest = out.split(":", 1)
            return f"{int(dom, 16):04x}:{rest.lower()}"
    except Exception:
        pass
    return None


def gpu_numa_node(device_index: int) -> int | None:
    addr = gpu_pci_address(device_index)
    if addr is None:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node: int):
    try:
        return _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except OSError:
        return set()


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node that hosts CUDA device ``device_index``.  Returns a small report
    (``bound``, ``node``, ``cpus``) for logs; never raises on missing topology information."""
    report = {"bound": False, "node": None, "cpus": None, "pci": gpu_pci_address(device_index)}
    node = gpu_numa_node(device_index)
    if node is None:
        report["why"] = "no NUMA node exposed for this GPU"
        return report
    report["node"] = node
    allowed = os.sched_getaffinity(0)
    cpus = node_cpus(node) & allowed
    if not cpus:
        report["why"] = "the node's CPUs are outside this process's cpuset"
        return report
    os.sched_setaffinity(0, cpus)
    report.update({"bound": True, "cpus": len(cpus)})
    return report
