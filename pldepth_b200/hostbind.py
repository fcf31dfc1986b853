"""Keep a rank's host side next to its GPU.

The end-to-end path (HostPipelinedStep) moves ~100 MB per step over PCIe from pinned host memory.  On a two-socket
box a rank whose pinned pages sit on the other socket pays the inter-socket link on every copy, and several ranks
doing so share that link.  Pinned pages are placed by first touch on the NUMA node of the allocating thread, so it
is enough to restrict the process to the CPUs that sysfs reports as local to the GPU *before* any pinned buffer
is allocated.  Nothing here is on the compute path; failure to bind is reported, never fatal.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_path(device_index):
    """sysfs directory of CUDA device ``device_index`` (None when it cannot be determined)."""
    import torch
    p = torch.cuda.get_device_properties(device_index)
    try:
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except AttributeError:
        # older torch: ask NVML for the device with the same UUID
        try:
            import pynvml
            pynvml.nvmlInit()
            want = str(p.uuid).replace("GPU-", "")
            bdf = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if u.replace("GPU-", "") == want:
                    b = pynvml.nvmlDeviceGetPciInfo(h).busId
                    b = b.decode() if isinstance(b, bytes) else b
                    bdf = b.lower()[-12:]          # NVML prints an 8-digit domain; sysfs uses 4
                    break
            if bdf is None:
                return None
        except Exception:
            return None
    path = "/sys/bus/pci/devices/" + bdf
    return path if os.path.isdir(path) else None


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs local to the GPU.  Returns a small report dict:
    ``{"bound": bool, "numa_node": int|None, "cpus": int, "why": str}``."""
    rep = {"bound": False, "numa_node": None, "cpus": 0, "why": ""}
    if not hasattr(os, "sched_setaffinity"):
        rep["why"] = "no sched_setaffinity"
        return rep
    try:
        path = gpu_pci_path(device_index)
        if path is None:
            rep["why"] = "no sysfs entry for the device"
            return rep
        with open(os.path.join(path, "numa_node")) as f:
            node = int(f.read().strip())
        rep["numa_node"] = node
        with open(os.path.join(path, "local_cpulist")) as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if node < 0 or not cpus or cpus == allowed:
            rep["why"] = "single node or nothing to narrow"
            rep["cpus"] = len(allowed)
            return rep
        os.sched_setaffinity(0, cpus)
        rep.update(bound=True, cpus=len(cpus))
    except Exception as e:  # containers may hide sysfs or forbid the call
        rep["why"] = "%s: %s" % (type(e).__name__, e)
    return rep
