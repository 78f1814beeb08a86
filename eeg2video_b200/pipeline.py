"""Host-resident recordings -> features: the end-to-end call (pinned host buffers in, pinned host features out).

Recordings are streamed through the GPU in chunks of whole blocks: the H2D copy of chunk i+1 (copy stream),
the fused kernel on chunk i (compute stream) and the D2H copy of chunk i-1's features (second copy stream)
overlap, so the end-to-end rate is the PCIe rate of the raw samples, not the sum of the three.
"""
import torch

from . import _lib, frontend, ops


class HostPipeline:
    """Reusable staging buffers + streams for `features_from_host`.

    n_ch, block_len: recording geometry; chunk_blocks: blocks per in-flight chunk (two chunks are staged).
    """

    def __init__(self, device, n_ch=62, block_len=104000, chunk_blocks=28, mode="500ms"):
        self.device = torch.device(device)
        self.mode = frontend._mode_id(mode)
        self.n_win = ops.WINDOWS_PER_CLIP[self.mode]
        self.n_ch, self.block_len, self.chunk_blocks = n_ch, block_len, chunk_blocks
        with torch.cuda.device(self.device):
            self.stage = [torch.empty((chunk_blocks, n_ch, block_len), dtype=torch.float32, device=self.device)
                          for _ in range(2)]
            self.h2d = torch.cuda.Stream()
            self.d2h = torch.cuda.Stream()
            self.compute = torch.cuda.Stream()

    def feature_shape(self, n_blocks):
        return (n_blocks * 200, self.n_win, self.n_ch, 5)

    def run(self, raw_host, de_host, psd_host):
        """raw_host: pinned float32 (n_blocks, n_ch, block_len); de_host / psd_host: pinned float32
        feature_shape(n_blocks).  Returns the accumulated status flags (int).  Synchronises before returning."""
        n_blocks = raw_host.shape[0]
        cb = self.chunk_blocks
        status_all = torch.zeros(1, dtype=torch.int32, device=self.device)
        staged = [None, None]          # events: stage[i] free again (its kernel finished)
        with torch.cuda.device(self.device):
            for i, lo in enumerate(range(0, n_blocks, cb)):
                hi = min(lo + cb, n_blocks)
                buf = self.stage[i & 1][: hi - lo]
                with torch.cuda.stream(self.h2d):
                    if staged[i & 1] is not None:
                        self.h2d.wait_event(staged[i & 1])
                    buf.copy_(raw_host[lo:hi], non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self.h2d)
                with torch.cuda.stream(self.compute):
                    self.compute.wait_event(ready)
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


def features_from_host(raw_host, mode="500ms", chunk_blocks=28, device="cuda", check=True):
    """Convenience wrapper: (.., 62, T) host tensor/array -> (de, psd) host tensors in the reference layout."""
    raw = torch.as_tensor(raw_host)
    lead = raw.shape[:-2]
    flat = raw.reshape((-1,) + tuple(raw.shape[-2:]))
    if not flat.is_pinned():
        flat = flat.contiguous().pin_memory()
    pipe = HostPipeline(device, flat.shape[1], flat.shape[2], min(chunk_blocks, max(flat.shape[0], 1)), mode)
    shape = pipe.feature_shape(flat.shape[0])
    de = torch.empty(shape, dtype=torch.float32).pin_memory()
    psd = torch.empty(shape, dtype=torch.float32).pin_memory()
    status = pipe.run(flat, de, psd)
    if check and status & _lib.STATUS_ZERO_POWER:
        raise ValueError("math domain error")
    n_win = shape[1]
    out_shape = tuple(lead) + (40, 5) + ((n_win,) if n_win > 1 else ()) + (flat.shape[1], 5)
    return de.reshape(out_shape), psd.reshape(out_shape)
