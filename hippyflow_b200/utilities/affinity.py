"""Host-side placement: one process per GPU should run (and first-touch its pinned staging buffers) on the CPU cores of
the NUMA node the GPU hangs off, otherwise host -> device copies cross the inter-socket link and several ranks share one
memory controller.  The reference leaves placement to ``mpirun --bind-to``; there is no launcher here, so the process
binds itself from NVML's view of the topology."""
import os


def bind_to_gpu_numa_node(gpu_index):
    """Restrict this process to the CPUs NVML reports as local to GPU ``gpu_index`` (PCI order = CUDA order when
    CUDA_DEVICE_ORDER=PCI_BUS_ID, which torchrun/gpurun boxes use).  Returns the sorted list of CPUs bound to, or
    None when NVML or sched_setaffinity is unavailable or the answer is empty (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = sorted(64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
