"""Batched predict engine: raw scans -> (class mask, class counts) with the whole hot path on the GPU.

One image goes through  K1 resize+trim -> FCN-ResNet50 plan -> K3 upsample+argmax -> K5 region removal+counts
(models.py:191-203 and 247-332 of the reference).  The engine removes the reference's per-image host round trips:
all K1 launches of a batch are issued first and their [first,last) rows are read back with ONE synchronisation,
then the per-image network / mask kernels are issued back to back and the counts come back with one more.
``run_host`` is the public end-to-end entry (pinned host buffers in, host masks out); H2D copies, K1 and the D2H of
finished masks run on side streams and overlap the segmentation of the previous chunk."""
import numpy as np
import torch

from . import ops


class PredictEngine:
    def __init__(self, model, device='cuda:0', threshold=150, raw_size=4096):
        self.model = model
        self.device = torch.device(device)
        self.threshold = threshold
        self.raw_size = raw_size
        self.out_w = raw_size // 4
        self._pre_ws = None
        self._ccl_ws = None
        self._copy_stream = None

    # -- buffers ------------------------------------------------------------------------------------------------
    def _buffers(self, n):
        S, Wo = self.raw_size, self.out_w
        if getattr(self, '_n', 0) < n:
            dev = self.device
            self._proc = torch.empty((n, (S // 4) * Wo * 3), dtype=torch.uint8, device=dev)
            self._fl = torch.empty((n, 2), dtype=torch.int32, device=dev)
            self._masks = torch.empty((n, (S // 4) * Wo), dtype=torch.uint8, device=dev)
            self._counts = torch.empty((n, 3), dtype=torch.int32, device=dev)
            self._n = n
            lib = ops._lib.load()
            self._pre_ws = torch.empty(lib.nbc_preprocess_workspace_bytes(S, S), dtype=torch.uint8, device=dev)
            self._ccl_ws = torch.empty(lib.nbc_ccl_workspace_bytes(1, S // 4, Wo), dtype=torch.uint8, device=dev)

    def _preprocess_into(self, i, raw, bgr, bottom_up):
        import ctypes as C
        lib = ops._lib.load()
        S = self.raw_size
        ops._lib.check(lib.nbc_preprocess_4x_u8(C.c_void_p(raw.data_ptr()), S, S, S * 3, (1 if bgr else 0) | (2 if bottom_up else 0),
                                                C.c_void_p(self._proc[i].data_ptr()), C.c_void_p(self._fl[i].data_ptr()),
                                                C.c_void_p(self._pre_ws.data_ptr()), self._pre_ws.numel(),
                                                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       'nbc_preprocess_4x_u8')

    def _segment(self, i, rows, exclude_nodes):
        """processed image i (rows x out_w) -> mask + counts, all asynchronous."""
        import ctypes as C
        lib = ops._lib.load()
        Wo = self.out_w
        img = self._proc[i, :rows * Wo * 3].view(1, rows, Wo, 3)
        low = self.model.lowres_logits_u8(img)
        mask = self._masks[i, :rows * Wo].view(1, rows, Wo)
        ops.upsample_argmax(low, (rows, Wo), out=mask)
        ops._lib.check(lib.nbc_remove_small_zones(C.c_void_p(mask.data_ptr()), 1, rows, Wo, self.threshold,
                                                  1 if exclude_nodes else 0, C.c_void_p(self._counts[i].data_ptr()),
                                                  C.c_void_p(self._ccl_ws.data_ptr()), self._ccl_ws.numel(),
                                                  C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       'nbc_remove_small_zones')

    # -- streams ---------------------------------------------------------------------------------------------------
    def _streams(self, depth):
        if self._copy_stream is None:
            dev = self.device
            self._copy_stream = torch.cuda.Stream(dev)      # H2D of raw scans
            self._pre_stream = torch.cuda.Stream(dev)       # K1 (resize + trim)
            self._out_stream = torch.cuda.Stream(dev)       # D2H of masks
            self._stage = [torch.empty(self.raw_size * self.raw_size * 3, dtype=torch.uint8, device=dev) for _ in range(depth)]
            self._staged = [torch.cuda.Event() for _ in range(depth)]
            self._freed = [torch.cuda.Event() for _ in range(depth)]
            self._fl_host = None

    def _run(self, n, get_raw, masks_host, bgr, bottom_up, exclude_nodes, chunk=8, depth=4):
        """Software pipeline over chunks of images:
             copy stream : H2D of the raw scans (host path only), ``depth`` staging buffers
             pre stream  : K1 for chunk c+1 + async read-back of its [first,last) rows
             main stream : network + K3 + K5 for chunk c      out stream : D2H of finished masks
        so the PCIe transfers, the preprocessing and the one host read-back per chunk hide behind the segmentation."""
        dev = self.device
        Wo = self.out_w
        with torch.cuda.device(dev):
            self._buffers(n)
            self._streams(depth)
            if self._fl_host is None or self._fl_host.shape[0] < n:
                self._fl_host = torch.empty((n, 2), dtype=torch.int32).pin_memory()
            main = torch.cuda.current_stream(dev)
            start = torch.cuda.Event()
            start.record(main)
            self._copy_stream.wait_event(start)
            self._pre_stream.wait_event(start)
            self._out_stream.wait_event(start)
            chunks = [(a, min(a + chunk, n)) for a in range(0, n, chunk)]
            pre_done = [torch.cuda.Event() for _ in chunks]
            state = {'issued': 0}

            def issue_pre(ci):
                a, b = chunks[ci]
                for i in range(a, b):
                    raw = get_raw(i)
                    if not raw.is_cuda:
                        k = state['issued'] % depth
                        with torch.cuda.stream(self._copy_stream):
                            if state['issued'] >= depth:
                                self._copy_stream.wait_event(self._freed[k])
                            self._stage[k].copy_(raw, non_blocking=True)
                            self._staged[k].record(self._copy_stream)
                        self._pre_stream.wait_event(self._staged[k])
                        with torch.cuda.stream(self._pre_stream):
                            self._preprocess_into(i, self._stage[k], bgr, bottom_up)
                            self._freed[k].record(self._pre_stream)
                        state['issued'] += 1
                    else:
                        with torch.cuda.stream(self._pre_stream):
                            self._preprocess_into(i, raw, bgr, bottom_up)
                with torch.cuda.stream(self._pre_stream):
                    self._fl_host[a:b].copy_(self._fl[a:b], non_blocking=True)
                    pre_done[ci].record(self._pre_stream)

            rows = [0] * n
            issue_pre(0)
            for ci, (a, b) in enumerate(chunks):
                if ci + 1 < len(chunks):
                    issue_pre(ci + 1)
                pre_done[ci].synchronize()
                main.wait_event(pre_done[ci])
                for i in range(a, b):
                    rows[i] = int(self._fl_host[i, 1] - self._fl_host[i, 0])
                    self._segment(i, rows[i], exclude_nodes)
                    if masks_host is not None:
                        ev = torch.cuda.Event()
                        ev.record(main)
                        with torch.cuda.stream(self._out_stream):
                            self._out_stream.wait_event(ev)
                            masks_host[i][:rows[i] * Wo].copy_(self._masks[i, :rows[i] * Wo], non_blocking=True)
            main.wait_stream(self._out_stream)
        return rows

    # -- device-resident batch ---------------------------------------------------------------------------------------
    def run_device(self, raws, bgr=True, bottom_up=True, exclude_nodes=False):
        """raws: list of u8 CUDA tensors, each a raw_size x raw_size x 3 pixel array already in HBM.
        Returns (rows per image, counts int32 [n,3] CUDA tensor, masks buffer [n, raw_size/4 * out_w] CUDA);
        asynchronous on the current stream apart from one tiny read-back per chunk of 8 images."""
        n = len(raws)
        rows = self._run(n, lambda i: raws[i], None, bgr, bottom_up, exclude_nodes)
        return rows, self._counts[:n], self._masks

    # -- end to end from pinned host memory --------------------------------------------------------------------------
    def run_host(self, raws_host, masks_host=None, bgr=True, bottom_up=True, exclude_nodes=False):
        """raws_host: list of pinned u8 CPU tensors (raw pixel arrays).  Returns (rows list, counts numpy [n,3],
        masks_host); when ``masks_host`` (pinned u8 tensors) is given every mask is copied back too."""
        n = len(raws_host)
        rows = self._run(n, lambda i: raws_host[i], masks_host, bgr, bottom_up, exclude_nodes)
        with torch.cuda.device(self.device):
            counts = self._counts[:n].cpu().numpy()      # syncs the stream: everything above is done
        return rows, counts, masks_host
