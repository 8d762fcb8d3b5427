"""NUMA placement of the host side of the end-to-end path (streaming.HostWarpPipeline, bench.py's `e2e` arm).

Pinned host buffers live on the NUMA node of the thread that allocates them.  With one process per GPU on a multi-socket
host, every rank should pin from the node its GPU's PCIe root hangs off, otherwise all ranks share one socket's memory
controllers and one inter-socket link (round 1 measured 8 ranks at 0.18 of linear end to end).  `bind_to_gpu_node`
restricts the calling process to the CPUs of that node BEFORE the buffers are pinned; it is a no-op (and says so) when
the box exposes a single node or the information is not available."""
import glob
import os
from typing import Dict, List, Optional


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as fh:
            return fh.read().strip()
    except OSError:
        return None


def _parse_cpulist(text: str) -> List[int]:
    cpus = []
    for part in text.split(','):
        part = part.strip()
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def node_cpus() -> Dict[int, List[int]]:
    """NUMA node -> CPUs, from sysfs."""
    nodes = {}
    for d in glob.glob('/sys/devices/system/node/node[0-9]*'):
        txt = _read(os.path.join(d, 'cpulist'))
        if txt:
            nodes[int(os.path.basename(d)[4:])] = _parse_cpulist(txt)
    return nodes


def gpu_numa_nodes() -> Dict[int, Optional[int]]:
    """CUDA device index -> NUMA node of its PCI function (None when sysfs does not say)."""
    out = {}
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:       # noqa: BLE001
        return out
    for i in range(n):
        node = None
        try:
            import torch
            bus = torch.cuda.get_device_properties(i).pci_bus_id
            dom = getattr(torch.cuda.get_device_properties(i), 'pci_domain_id', 0)
            devid = getattr(torch.cuda.get_device_properties(i), 'pci_device_id', 0)
            txt = _read('/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, devid))
            if txt is not None and int(txt) >= 0:
                node = int(txt)
        except Exception:   # noqa: BLE001
            node = None
        out[i] = node
    return out


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node of CUDA device `device_index` (intersection with the CPUs it
    is allowed to use).  Returns what was done: {'node', 'cpus', 'bound', 'why'}."""
    nodes = node_cpus()
    info = {'node': None, 'cpus': None, 'bound': False, 'why': ''}
    if len(nodes) <= 1:
        info['why'] = 'single NUMA node visible (%s)' % (sorted(nodes) or 'none')
        return info
    node = gpu_numa_nodes().get(device_index)
    if node is None or node not in nodes:
        info['why'] = 'NUMA node of GPU %d unknown' % device_index
        return info
    allowed = os.sched_getaffinity(0)
    cpus = sorted(set(nodes[node]) & allowed)
    info['node'] = node
    if not cpus:
        info['why'] = 'no allowed CPU on node %d' % node
        return info
    os.sched_setaffinity(0, cpus)
    info.update(cpus='%d-%d (%d)' % (cpus[0], cpus[-1], len(cpus)), bound=True, why='bound before pinning')
    return info
