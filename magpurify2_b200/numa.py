"""Host placement for multi-GPU runs: keep a rank's (or a device thread's) staging memory and packer threads on
the NUMA node its GPU hangs off.

Eight ranks each pushing their shard through one host saturate the inter-socket link long before PCIe when the
pinned buffers of all of them sit on one node (round 1: end-to-end efficiency 0.52 at 4 GPUs, 0.41 at 8).  Linux
places freshly touched pages on the node of the touching CPU, so pinning the calling thread to the CPUs of the
GPU's node BEFORE the pinned buffers are allocated is all it takes; no libnuma needed.  Everything here is best
effort: inside a container without NUMA visibility it reports that and changes nothing."""
import os
import subprocess


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device):
    """(pci_bus_id, numa_node or None) of CUDA device `device` (torch is used for the PCI id only)."""
    import torch

    props = torch.cuda.get_device_properties(device)
    bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
    node = _read(f"/sys/bus/pci/devices/{bus}/numa_node")
    try:
        node = int(node)
    except (TypeError, ValueError):
        node = None
    return bus, (node if node is not None and node >= 0 else None)


def nvml_cpu_affinity(device):
    """CPUs the driver calls local to the GPU (NVML: what `nvidia-smi topo -m` prints as CPU Affinity), or None."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(device)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (getattr(bus, "pci_domain_id", 0), bus.pci_bus_id,
                                                                         bus.pci_device_id)).encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        return {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_node(device):
    """Restricts the CALLING THREAD to the CPUs of the GPU's NUMA node (intersected with what it may use now).
    Returns a small record for logs / the bench line."""
    info = {"device": int(device), "bound": False}
    try:
        bus, node = gpu_numa_node(device)
        info.update(pci=bus, numa_node=node)
        if node is None:
            node_cpus = nvml_cpu_affinity(device)  # sysfs is hidden inside some containers; the driver still knows
            info["source"] = "nvml"
            if not node_cpus:
                info["why"] = "no NUMA node reported for the device (single node, or hidden by the container)"
                return info
        else:
            node_cpus = _parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist"))
        allowed = os.sched_getaffinity(0)
        use = node_cpus & allowed
        info["node_cpus"] = len(node_cpus)
        if not use or use == allowed:
            info["why"] = "the node's CPUs are all (or none) of the CPUs this process may use"
            return info
        os.sched_setaffinity(0, use)
        info.update(bound=True, cpus=len(use))
    except Exception as ex:  # noqa: BLE001 -- placement is an optimisation, never a failure
        info["why"] = f"{type(ex).__name__}: {ex}"
    return info


def topology_text():
    """`nvidia-smi topo -m` (GPU <-> GPU / NIC / CPU affinity matrix) for the record; '' when unavailable."""
    try:
        return subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception:  # noqa: BLE001
        return ""
