"""Host-resident recordings -> features: the end-to-end call (pinned host buffers in, pinned host features out).

Recordings are streamed through the GPU in chunks of whole blocks: the H2D copy of chunk i+1 (copy stream),
the fused kernel on chunk i (compute stream) and the D2H copy of chunk i-1's features (second copy stream)
overlap, so the end-to-end rate is the PCIe rate of the samples that have to cross, not the sum of the three.

Only the live samples need to cross PCIe: a SEED-DV block is 40 concepts x (600 hint + 2000 clip) samples, so each
channel row can be uploaded with ONE strided DMA (cudaMemcpy2DAsync: 40 pieces of 8000 B out of every 10400 B) into a
compact (blocks, channels, 40, 2000) staging tensor -- 23 % fewer bytes than the raw row -- which the kernel reads
directly (eegfe_de_psd_from_concepts).  How fast a host serves such a strided read differs from box to box (measured on
B200 hosts of one pool: 51 GB/s on most, 29 GB/s on some, against 53-55 GB/s for a plain contiguous copy), so by
default the pipeline times both layouts on its first chunk and keeps the faster one (compact="auto").
"""
import torch

from . import _lib, frontend, ops

CONCEPTS, HINT, CLIPS_LEN, CONCEPT_LEN = 40, 600, 2000, 2600


def bind_to_gpu_numa_node(device):
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs), so that the pinned host
    buffers it allocates next are node-local to the GPU's PCIe root.  Returns the CPU set used, or None when the
    platform does not say (VMs often report numa_node = -1) -- in which case nothing is changed."""
    import os
    try:
        bus = torch.cuda.get_device_properties(torch.device(device)).pci_bus_id
        dom = torch.cuda.get_device_properties(torch.device(device)).pci_domain_id
        dev = torch.cuda.get_device_properties(torch.device(device)).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(os.path.join(path, "numa_node")) as f:
            if int(f.read().strip()) < 0:
                return None
        with open(os.path.join(path, "local_cpulist")) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except (OSError, ValueError, AttributeError):
        return None


class HostPipeline:
    """Reusable staging buffers + streams for `features_from_host`.

    n_ch, block_len: recording geometry; chunk_blocks: blocks per in-flight chunk (two chunks are staged).
    compact: True -- upload only the 2000 live samples of every 2600 (strided DMA); False -- whole rows (contiguous DMA,
    23 % more bytes); "auto" (default) -- time both on the first chunk of the first run() and keep the faster
    (`self.compact` then says which, `self.upload_probe` holds the two times; decided once per device and geometry).
    """

    _layout_cache = {}          # (device index, n_ch, block_len, chunk_blocks) -> (compact, probe)

    def __init__(self, device, n_ch=62, block_len=104000, chunk_blocks=28, mode="500ms", compact="auto"):
        self.device = torch.device(device)
        self.mode = frontend._mode_id(mode)
        self.n_win = ops.WINDOWS_PER_CLIP[self.mode]
        self.n_ch, self.block_len, self.chunk_blocks = n_ch, block_len, chunk_blocks
        if block_len < CONCEPTS * CONCEPT_LEN:
            raise RuntimeError("Segment length mismatch")
        self.upload_probe = None
        self.compact = None if compact == "auto" else bool(compact)
        with torch.cuda.device(self.device):
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self._cache_key = (idx, n_ch, block_len, chunk_blocks)
            if self.compact is None and self._cache_key in HostPipeline._layout_cache:
                self.compact, self.upload_probe = HostPipeline._layout_cache[self._cache_key]
            # one flat buffer per stage, big enough for either layout; the layout in use is a view of it
            per_row = block_len if self.compact in (None, False) else CONCEPTS * CLIPS_LEN
            self._flat = [torch.empty(chunk_blocks * n_ch * per_row, dtype=torch.float32, device=self.device)
                          for _ in range(2)]
            self.h2d = torch.cuda.Stream()
            self.d2h = torch.cuda.Stream()
            self.compute = torch.cuda.Stream()

    def _stage(self, i, n_blocks, compact):
        shape = (n_blocks, self.n_ch, CONCEPTS, CLIPS_LEN) if compact else (n_blocks, self.n_ch, self.block_len)
        n = n_blocks * self.n_ch * (CONCEPTS * CLIPS_LEN if compact else self.block_len)
        return self._flat[i][:n].view(shape)

    def _choose_layout(self, raw_host):
        """Time the strided and the contiguous upload of the first chunk (CUDA events on the copy stream, second of two
        runs each) and keep the layout that takes less time."""
        hi = min(self.chunk_blocks, raw_host.shape[0])
        times = {}
        with torch.cuda.stream(self.h2d):
            for compact in (True, False):
                self.compact = compact
                buf = self._stage(0, hi, compact)
                for _ in range(2):
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record(self.h2d)
                    self._upload(buf, raw_host, 0, hi)
                    t1.record(self.h2d)
                    self.h2d.synchronize()
                times[compact] = t0.elapsed_time(t1)
        self.compact = times[True] <= times[False]
        self.upload_probe = {"strided_live_samples_ms": times[True], "contiguous_rows_ms": times[False],
                             "blocks": int(hi)}
        HostPipeline._layout_cache[self._cache_key] = (self.compact, self.upload_probe)

    def feature_shape(self, n_blocks):
        return (n_blocks * 200, self.n_win, self.n_ch, 5)

    def h2d_bytes(self, n_blocks):
        """bytes run() uploads for n_blocks blocks (after the layout has been chosen)."""
        per_row = CONCEPTS * CLIPS_LEN if self.compact else self.block_len
        return n_blocks * self.n_ch * per_row * 4

    def _upload(self, buf, raw_host, lo, hi):
        if not self.compact:
            buf.copy_(raw_host[lo:hi], non_blocking=True)
            return
        lib = _lib.load()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.block_len == CONCEPTS * CONCEPT_LEN:
            # rows are back to back: the whole chunk is one uniform 2-D pattern
            src = raw_host.data_ptr() + (lo * self.n_ch * self.block_len + HINT) * 4
            _lib.check(lib.eegfe_copy2d_async(buf.data_ptr(), CLIPS_LEN * 4, src, CONCEPT_LEN * 4, CLIPS_LEN * 4,
                                              (hi - lo) * self.n_ch * CONCEPTS, 1, stream))
        else:
            for b in range(lo, hi):
                for ch in range(self.n_ch):
                    src = raw_host.data_ptr() + ((b * self.n_ch + ch) * self.block_len + HINT) * 4
                    dst = buf.data_ptr() + ((b - lo) * self.n_ch + ch) * CONCEPTS * CLIPS_LEN * 4
                    _lib.check(lib.eegfe_copy2d_async(dst, CLIPS_LEN * 4, src, CONCEPT_LEN * 4, CLIPS_LEN * 4,
                                                      CONCEPTS, 1, stream))

    def run(self, raw_host, de_host, psd_host):
        """raw_host: pinned, contiguous float32 (n_blocks, n_ch, block_len); de_host / psd_host: pinned float32
        feature_shape(n_blocks).  Returns the accumulated status flags (int).  Synchronises before returning."""
        if not (raw_host.is_contiguous() and raw_host.dtype == torch.float32):
            raise ValueError("raw_host must be contiguous float32 (features_from_host converts other types)")
        n_blocks = raw_host.shape[0]
        cb = self.chunk_blocks
        if self.compact is None:
            if n_blocks == 0:
                self.compact = True
            else:
                with torch.cuda.device(self.device):
                    self._choose_layout(raw_host)
        staged = [None, None]          # events: stage[i] free again (its kernel finished)
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self.compute):      # zero-filled on the stream that ORs into it (no cross-stream race)
                status_all = torch.zeros(1, dtype=torch.int32, device=self.device)
            for i, lo in enumerate(range(0, n_blocks, cb)):
                hi = min(lo + cb, n_blocks)
                buf = self._stage(i & 1, hi - lo, self.compact)
                with torch.cuda.stream(self.h2d):
                    if staged[i & 1] is not None:
                        self.h2d.wait_event(staged[i & 1])
                    self._upload(buf, raw_host, lo, hi)
                    ready = torch.cuda.Event()
                    ready.record(self.h2d)
                with torch.cuda.stream(self.compute):
                    self.compute.wait_event(ready)
                    if self.compact:
                        de, psd, status = ops.de_psd_from_concepts(buf, self.mode)
                    else:
                        de, psd, status = ops.de_psd_from_raw(buf, self.mode)
                    status_all |= status
                    done = torch.cuda.Event()
                    done.record(self.compute)
                    staged[i & 1] = done
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    de_host[lo * 200:hi * 200].copy_(de, non_blocking=True)
                    psd_host[lo * 200:hi * 200].copy_(psd, non_blocking=True)
                    de.record_stream(self.d2h)
                    psd.record_stream(self.d2h)
            self.d2h.synchronize()
            self.compute.synchronize()
            return int(status_all.item())


def features_from_host(raw_host, mode="500ms", chunk_blocks=28, device="cuda", check=True, compact="auto"):
    """Convenience wrapper: (.., 62, T) host tensor/array -> (de, psd) host tensors in the reference layout."""
    raw = torch.as_tensor(raw_host)
    lead = raw.shape[:-2]
    flat = raw.reshape((-1,) + tuple(raw.shape[-2:]))
    if flat.dtype != torch.float32:
        # other element types (float64 recordings, int16 codes): block by block through the device, which rounds to
        # float32 far faster than a host pass; the float32 copy is what the pipeline then streams
        dev = torch.device(device)
        out = torch.empty(flat.shape, dtype=torch.float32).pin_memory()
        for b in range(flat.shape[0]):
            out[b].copy_(flat[b].to(dev, non_blocking=False).to(torch.float32))
        flat = out
    if not flat.is_pinned():
        flat = flat.contiguous().pin_memory()
    pipe = HostPipeline(device, flat.shape[1], flat.shape[2], min(chunk_blocks, max(flat.shape[0], 1)), mode, compact)
    shape = pipe.feature_shape(flat.shape[0])
    de = torch.empty(shape, dtype=torch.float32).pin_memory()
    psd = torch.empty(shape, dtype=torch.float32).pin_memory()
    status = pipe.run(flat, de, psd)
    if check and status & _lib.STATUS_ZERO_POWER:
        raise ValueError("math domain error")
    n_win = shape[1]
    out_shape = tuple(lead) + (40, 5) + ((n_win,) if n_win > 1 else ()) + (flat.shape[1], 5)
    return de.reshape(out_shape), psd.reshape(out_shape)
