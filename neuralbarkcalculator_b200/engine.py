"""Batched predict engine: raw scans -> (class mask, class counts) with the whole hot path on the GPU.

One image goes through  K1 resize+trim -> FCN-ResNet50 plan -> K3 upsample+argmax -> K5 region removal+counts
(models.py:191-203 and 247-332 of the reference).  The engine removes the reference's per-image host round trips
entirely: images are processed in RAGGED batches (a canvas of ``chunk`` images of different trimmed heights; every
kernel reads the per-image heights K1 produced on the device), so nothing is read back until the final counts.
``run_host`` is the public end-to-end entry (pinned host buffers in, host masks out); H2D copies, K1 and the D2H of
finished masks run on side streams and overlap the segmentation of the previous chunk."""
import ctypes as C

import numpy as np
import torch

from . import ops


class PredictEngine:
    """chunk: images per ragged batch (one network launch sequence per chunk)."""

    def __init__(self, model, device='cuda:0', threshold=150, raw_size=4096, chunk=8, depth=4):
        self.model = model
        self.device = torch.device(device)
        self.threshold = threshold
        self.raw_size = raw_size
        self.out_w = raw_size // 4
        self.chunk = chunk
        self.depth = depth
        self._n = 0
        self._copy_stream = None

    # -- buffers ------------------------------------------------------------------------------------------------
    def _buffers(self, n):
        S, Wo = self.raw_size, self.out_w
        Hc = S // 4
        if self._n < n:
            dev = self.device
            self._proc = torch.empty((n, Hc, Wo, 3), dtype=torch.uint8, device=dev)      # canvas: image i in rows [0, h_i)
            self._fl = torch.empty((n, 2), dtype=torch.int32, device=dev)
            self._heights = torch.empty(n, dtype=torch.int32, device=dev)
            self._masks = torch.empty((n, Hc, Wo), dtype=torch.uint8, device=dev)
            self._counts = torch.empty((n, 3), dtype=torch.int32, device=dev)
            self._n = n
            lib = ops._lib.load()
            self._pre_ws = torch.empty(lib.nbc_preprocess_workspace_bytes(S, S), dtype=torch.uint8, device=dev)
            self._ccl_ws = torch.empty(lib.nbc_ccl_workspace_bytes(self.chunk, Hc, Wo), dtype=torch.uint8, device=dev)
            hl = (((Hc - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1
            wl = (((Wo - 1) // 2 + 1 - 1) // 2 + 1 - 1) // 2 + 1
            self._logits = [torch.empty((self.chunk, 3, hl, wl), dtype=torch.float32, device=dev) for _ in range(2)]
            self._logits_free = [None, None]
            self._fl_host = torch.empty((n, 2), dtype=torch.int32).pin_memory()

    def _streams(self):
        if self._copy_stream is None:
            dev = self.device
            self._copy_stream = torch.cuda.Stream(dev)      # H2D of raw scans
            self._pre_stream = torch.cuda.Stream(dev)       # K1 (resize + trim)
            self._out_stream = torch.cuda.Stream(dev)       # D2H of masks
            self._post_stream = torch.cuda.Stream(dev)      # K3 + K5 of the previous chunk, under the next network pass
            self._stage = [torch.empty(self.raw_size * self.raw_size * 3, dtype=torch.uint8, device=dev)
                           for _ in range(self.depth)]
            self._staged = [torch.cuda.Event() for _ in range(self.depth)]
            self._freed = [torch.cuda.Event() for _ in range(self.depth)]

    def _preprocess_into(self, i, raw, bgr, bottom_up):
        lib = ops._lib.load()
        S = self.raw_size
        ops._lib.check(lib.nbc_preprocess_4x_u8(C.c_void_p(raw.data_ptr()), S, S, S * 3, (1 if bgr else 0) | (2 if bottom_up else 0),
                                                C.c_void_p(self._proc[i].data_ptr()), C.c_void_p(self._fl[i].data_ptr()),
                                                C.c_void_p(self._pre_ws.data_ptr()), self._pre_ws.numel(),
                                                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                       'nbc_preprocess_4x_u8')

    def _segment_chunk(self, ci, a, b, exclude_nodes):
        """Images a..b-1 as ONE ragged batch: network, K3 and K5 read the per-image heights on the device.  The network
        runs on the main stream; K3 (upsample+argmax) and K5 (region removal + counts) follow on the post stream, so
        they overlap the network pass of the next chunk (two logits buffers).  Returns the event that marks the chunk's
        masks and counts as final."""
        plan = self.model.native_plan()
        main = torch.cuda.current_stream(self.device)
        k = ci & 1
        if self._logits_free[k] is not None:
            main.wait_event(self._logits_free[k])        # K3 of chunk ci-2 has consumed this logits buffer
        logits = plan.forward_ragged(self._proc[a:b], heights=self._heights[a:b], out=self._logits[k][:b - a])
        net_done = torch.cuda.Event()
        net_done.record(main)
        done = torch.cuda.Event()
        with torch.cuda.stream(self._post_stream):
            self._post_stream.wait_event(net_done)
            ops.upsample_argmax_ragged(logits, self._heights[a:b], (self.raw_size // 4, self.out_w), out=self._masks[a:b])
            free = torch.cuda.Event()
            free.record(self._post_stream)
            self._logits_free[k] = free
            ops.remove_small_zones_ragged(self._masks[a:b], self._heights[a:b], self.threshold, exclude_nodes,
                                          workspace=self._ccl_ws, counts=self._counts[a:b])
            done.record(self._post_stream)
        return done

    def _run(self, n, get_raw, masks_host, bgr, bottom_up, exclude_nodes):
        """Software pipeline over chunks of images, with NO host synchronisation inside:
             copy stream : H2D of the raw scans (host path only), ``depth`` staging buffers
             pre stream  : K1 + heights for chunk c+1
             main stream : ragged network for chunk c                post stream: K3 + K5 for chunk c-1
             out stream  : D2H of finished masks"""
        dev = self.device
        chunk, depth = self.chunk, self.depth
        with torch.cuda.device(dev):
            self._buffers(n)
            self._streams()
            main = torch.cuda.current_stream(dev)
            start = torch.cuda.Event()
            start.record(main)
            for st in (self._copy_stream, self._pre_stream, self._out_stream, self._post_stream):
                st.wait_event(start)
            self._logits_free = [None, None]
            chunks = [(a, min(a + chunk, n)) for a in range(0, n, chunk)]
            pre_done = [torch.cuda.Event() for _ in chunks]
            issued = 0
            for ci, (a, b) in enumerate(chunks):
                for i in range(a, b):
                    raw = get_raw(i)
                    if not raw.is_cuda:
                        k = issued % depth
                        with torch.cuda.stream(self._copy_stream):
                            if issued >= depth:
                                self._copy_stream.wait_event(self._freed[k])
                            self._stage[k].copy_(raw, non_blocking=True)
                            self._staged[k].record(self._copy_stream)
                        self._pre_stream.wait_event(self._staged[k])
                        with torch.cuda.stream(self._pre_stream):
                            self._preprocess_into(i, self._stage[k], bgr, bottom_up)
                            self._freed[k].record(self._pre_stream)
                        issued += 1
                    else:
                        with torch.cuda.stream(self._pre_stream):
                            self._preprocess_into(i, raw, bgr, bottom_up)
                with torch.cuda.stream(self._pre_stream):
                    ops.heights_from_first_last(self._fl[a:b], out=self._heights[a:b])
                    pre_done[ci].record(self._pre_stream)
                main.wait_event(pre_done[ci])
                ev = self._segment_chunk(ci, a, b, exclude_nodes)
                if masks_host is not None:
                    with torch.cuda.stream(self._out_stream):
                        self._out_stream.wait_event(ev)
                        for i in range(a, b):
                            masks_host[i].copy_(self._masks[i].view(-1), non_blocking=True)
            with torch.cuda.stream(self._pre_stream):
                self._fl_host[:n].copy_(self._fl[:n], non_blocking=True)
            main.wait_stream(self._pre_stream)
            main.wait_stream(self._post_stream)
            main.wait_stream(self._out_stream)

    def rows(self, n):
        """Valid rows per image of the last batch (host list); synchronises."""
        torch.cuda.current_stream(self.device).synchronize()
        fl = self._fl_host[:n]
        return (fl[:, 1] - fl[:, 0]).tolist()

    # -- device-resident batch ---------------------------------------------------------------------------------------
    def run_device(self, raws, bgr=True, bottom_up=True, exclude_nodes=False):
        """raws: list of u8 CUDA tensors, each a raw_size x raw_size x 3 pixel array already in HBM.  Fully
        asynchronous.  Returns (counts int32 [n,3] CUDA, mask canvas u8 [n, raw/4, raw/4] CUDA, heights int32 [n] CUDA):
        image i's mask is ``masks[i, :heights[i]]``."""
        n = len(raws)
        self._run(n, lambda i: raws[i], None, bgr, bottom_up, exclude_nodes)
        return self._counts[:n], self._masks[:n], self._heights[:n]

    # -- end to end from pinned host memory --------------------------------------------------------------------------
    def run_host(self, raws_host, masks_host=None, bgr=True, bottom_up=True, exclude_nodes=False):
        """raws_host: list of pinned u8 CPU tensors (raw pixel arrays).  Returns (rows list, counts numpy [n,3],
        masks_host): when ``masks_host`` (pinned u8 tensors of raw/4 * raw/4 bytes) is given every mask canvas is copied
        back; image i's mask is the first rows[i] * (raw/4) bytes."""
        n = len(raws_host)
        self._run(n, lambda i: raws_host[i], masks_host, bgr, bottom_up, exclude_nodes)
        with torch.cuda.device(self.device):
            counts = self._counts[:n].cpu().numpy()      # syncs the stream: everything above is done
        return self.rows(n), counts, masks_host
