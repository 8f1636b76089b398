"""Host-side plumbing between the nn.Module mirror of the reference and the C ABI of libbarvae.so.

* ``Act``        -- an NHWC bf16/fp32 activation view (tensor, geometry, channel pitch/offset).
* ``GemmLayer``  -- one Conv2d / ConvTranspose2d / Linear weight: plans the implicit-GEMM phases (sub-pixel
                    decomposition of strided transposed convolutions), keeps the packed bf16 operands in sync with
                    the fp32 master parameter, and issues forward / data-gradient / weight-gradient calls.
* ``NormBlock``  -- InstanceNorm(+CBAM)(+residual)+activation through bvae_nb_forward / bvae_nb_backward.
* ``FlatParams`` -- re-homes a module's parameters and gradients into flat fp32 buckets (one Adam launch, one
                    NCCL all-reduce per bucket).

PyTorch here is device memory + streams only; all arithmetic on the hot path happens inside libbarvae.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, NbDesc, WgradDesc

BF16 = torch.bfloat16
_PARAM_EPOCH = [0]          # bumped whenever libbarvae itself rewrites parameters (fused Adam)
_IMPL = [_lib.IMPL_AUTO]    # contraction implementation selector (tests flip it to compare SIMT vs tcgen05)
_RAW_F32 = [True]           # dtype of raw conv outputs feeding an InstanceNorm (see DESIGN.md "Parity tolerances")
_FUSE_STATS = [os.environ.get("BVAE_FUSE_STATS", "0") == "1"]   # opt-in: let the conv epilogue accumulate the InstanceNorm statistics of its output
                            # (measured: the extra epilogue work costs the small-channel layers what the saved
                            # statistics pass gains -- 56.6 vs 54.4 ms/step -- so it is off by default)


def set_impl(impl: int):
    _IMPL[0] = impl


def get_impl() -> int:
    return _IMPL[0]


def set_raw_f32(flag: bool):
    _RAW_F32[0] = bool(flag)


def raw_dtype():
    return torch.float32 if _RAW_F32[0] else BF16


def set_fuse_stats(flag: bool):
    _FUSE_STATS[0] = bool(flag)


_SAVE_FOR_BACKWARD = True


class saving:
    """``with saving(False):`` -- forward passes issued inside do not write what only the backward needs (the bf16
    normalised activation of every norm site: 2 of the 16 bytes per element the norm-block forward moves).  Entered by
    the trunk / decoder autograd nodes with ``any(ctx.needs_input_grad)``: sampling (maker_bar.py:32-44, torch.no_grad)
    and frozen-generator passes save nothing."""

    def __init__(self, flag: bool):
        self.flag = bool(flag)

    def __enter__(self):
        global _SAVE_FOR_BACKWARD
        self.prev, _SAVE_FOR_BACKWARD = _SAVE_FOR_BACKWARD, self.flag

    def __exit__(self, *exc):
        global _SAVE_FOR_BACKWARD
        _SAVE_FOR_BACKWARD = self.prev
        return False


def bump_param_epoch():
    _PARAM_EPOCH[0] += 1


# ---- weight gradients on a companion stream ------------------------------------------------------------------
# Nothing in a backward pass waits for a weight gradient until the optimiser: inside `async_wgrad()` every
# GemmLayer.wgrad launch goes to a stream paired with the current one, so the tensor-bound weight-gradient kernels run
# next to the HBM-bound norm-block sweeps of the data-gradient chain.  Leaving the context makes the current stream wait
# for the companion.  BVAE_WGRAD_STREAM=0 disables.
_WGRAD_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}
_WGRAD_ASYNC = [0]
# Operands of in-flight companion-stream launches are kept referenced until the join instead of record_stream()-ed:
# record_stream defers the caching allocator's reuse of a block until an event query succeeds, and with the host
# several steps ahead of the GPU those deferred frees pile up until the allocator falls back to cudaMalloc/cudaFree
# (measured: sporadic 55-116 ms steps).  After the join every later use on the current stream is ordered anyway.
_WGRAD_KEEP: Dict[tuple, list] = {}


def _wgrad_pair():
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    ws = _WGRAD_STREAMS.get(key)
    if ws is None:
        ws = _WGRAD_STREAMS[key] = torch.cuda.Stream(device=cur.device)
    return cur, ws


class async_wgrad:
    def __enter__(self):
        self.on = os.environ.get("BVAE_WGRAD_STREAM", "1") != "0"
        if self.on:
            _WGRAD_ASYNC[0] += 1
        return self

    def __exit__(self, *exc):
        if self.on:
            _WGRAD_ASYNC[0] -= 1
            cur, ws = _wgrad_pair()
            cur.wait_stream(ws)
            _WGRAD_KEEP.pop((cur.device.index, cur.cuda_stream), None)
        return False


# ---- independent branches of one block on a companion stream ----------------------------------------------------
# The two transposed-convolution branches of every decoder up-sampling block, and the decoder's two head stems, are
# independent until they are concatenated; on one stream their latency-bound small-map norm blocks run back to back
# with most SMs idle.  `with fork() as f: ...; f.join()` issues the enclosed launches on a companion of the current
# stream.  Rules kept by the callers: no weight-gradient launch inside a fork (async_wgrad joins only the companion of
# the stream that is current at its exit); tensors allocated inside belong to the companion's pool and are reused
# there only after the next fork's wait on the current stream.  Checked on the device in round 2
# (tests/test_gpu_model.py::test_decoder_branch_streams_agree: same parameters as the one-stream order within the
# run-to-run floor) and measured: 41.73 -> 41.37 ms per 512-bar step.  BVAE_DEC_STREAMS=0 disables (and BVAE_STREAMS=0
# still switches every overlap off).
_FORK_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}


def fork_enabled() -> bool:
    return os.environ.get("BVAE_DEC_STREAMS", "1") == "1" and os.environ.get("BVAE_STREAMS", "1") != "0"


class fork:
    def __init__(self):
        self.on = fork_enabled()

    def __enter__(self):
        if self.on:
            self.cur = torch.cuda.current_stream()
            key = (self.cur.device.index, self.cur.cuda_stream)
            self.side = _FORK_STREAMS.get(key)
            if self.side is None:
                self.side = _FORK_STREAMS[key] = torch.cuda.Stream(device=self.cur.device)
            self.side.wait_stream(self.cur)
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self._ctx.__exit__(*exc)
        return False

    def join(self):
        if self.on:
            self.cur.wait_stream(self.side)


# ---- live per-kernel-class timing with CUDA events (bench.py's roofline block) -----------------------------
_PROF = [None]


def profile_begin():
    _PROF[0] = {"_start": torch.cuda.Event(enable_timing=True)}
    _PROF[0]["_start"].record()


def _timed(tag, fn, *args, detail=None, flops=0.0, cls=None):
    """run one library call; when profiling, bracket it with CUDA events on the launching stream.  ``flops`` /
    ``cls`` (contraction launches only): algorithmic FLOPs of the launch (SURVEY.md section 8 convention) and its
    channel class, for the per-class tensor roofline in bench.py."""
    prof = _PROF[0]
    if prof is None:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    kern = _lib.last_kernel() if cls is not None else None      # contraction launches: the kernel instance that ran
    prof.setdefault(tag, []).append((e0, e1, detail, flops, cls, kern))
    return rc


def profiling() -> bool:
    return _PROF[0] is not None


def profile_end() -> dict:
    prof, _PROF[0] = _PROF[0], None
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    out = {"total_ms": prof.pop("_start").elapsed_time(end)}
    detail = {}
    classes = {}
    kernels = {}
    for tag, evs in prof.items():
        out[tag] = sum(ev[0].elapsed_time(ev[1]) for ev in evs)
        out["n_" + tag] = len(evs)
        for a, b, d, fl, cl, kern in evs:
            ms = a.elapsed_time(b)
            if d is not None:
                k = tag + ":" + d
                t, n = detail.get(k, (0.0, 0))
                detail[k] = (t + ms, n + 1)
            if cl is not None:
                c = classes.setdefault(cl, [0.0, 0.0, 0])      # ms, algorithmic FLOPs, launches
                c[0] += ms
                c[1] += fl
                c[2] += 1
            if kern:
                kk = kernels.setdefault(kern, {"ms": 0.0, "flops": 0.0, "launches": 0, "layers": {}})
                kk["ms"] += ms
                kk["flops"] += fl
                kk["launches"] += 1
                if d is not None:
                    lay = kk["layers"].setdefault(d, [0.0, 0.0, 0])
                    lay[0] += ms
                    lay[1] += fl
                    lay[2] += 1
    out["detail"] = detail
    out["classes"] = classes
    out["kernels"] = kernels
    out["n_gemm"] = out.get("n_conv_gemm", 0) + out.get("n_wgrad_gemm", 0)
    return out


class Act:
    """NHWC activation view: element (n,h,w,c) lives at t.data_ptr() + (((n*H+h)*W+w)*pitch + off + c)*esize."""
    __slots__ = ("t", "N", "H", "W", "C", "pitch", "off")

    def __init__(self, t: torch.Tensor, N: int, H: int, W: int, C: int, pitch: Optional[int] = None, off: int = 0):
        self.t, self.N, self.H, self.W, self.C = t, N, H, W, C
        self.pitch = C if pitch is None else pitch
        self.off = off

    @staticmethod
    def empty(N, H, W, C, dtype=BF16, device="cuda") -> "Act":
        return Act(torch.empty((N, H, W, C), dtype=dtype, device=device), N, H, W, C)

    def slice(self, off: int, C: int) -> "Act":
        return Act(self.t, self.N, self.H, self.W, C, self.pitch, self.off + off)

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + self.off * self.t.element_size()

    @property
    def f32(self) -> bool:
        return self.t.dtype == torch.float32

    def dense(self) -> torch.Tensor:
        """[N,H,W,C] tensor view (strided if this is a slice of a concatenated buffer)."""
        v = self.t.view(self.N, self.H, self.W, self.pitch)
        return v[..., self.off:self.off + self.C]


def grad_ptr(p: Optional[torch.Tensor]) -> int:
    """Device pointer of p.grad, creating a zero gradient if needed.  Frozen parameters (requires_grad=False,
    agent/barGen.py:143-149) get a throw-away sink so kernels can still accumulate somewhere."""
    if p is None:
        return 0
    if not p.requires_grad:
        sink = getattr(p, "_bvae_sink", None)
        if sink is None or sink.shape != p.shape or sink.device != p.device:
            sink = torch.zeros_like(p)
            p._bvae_sink = sink
        return sink.data_ptr()
    if p.grad is None:
        flat = getattr(p, "_bvae_flat", None)
        if flat is not None:
            p.grad = flat.grad_view(p).zero_()
        else:
            p.grad = torch.zeros_like(p)
    return p.grad.data_ptr()


# ---------------------------------------------------------------------------------------------------------
# implicit-GEMM layers
# ---------------------------------------------------------------------------------------------------------
class _Phase:
    __slots__ = ("taps", "dy", "dx", "sy", "sx", "osy", "osx", "ooy", "oox", "k_off")

    def __init__(self, taps, dy, dx, sy, sx, osy, osx, ooy, oox, k_off):
        self.taps, self.dy, self.dx = taps, dy, dx
        self.sy, self.sx, self.osy, self.osx, self.ooy, self.oox = sy, sx, osy, osx, ooy, oox
        self.k_off = k_off          # first packed tap of this phase inside the packed weight rows


def _transposed_phases(kh, kw, sy, sx, py, px):
    """Sub-pixel decomposition of out[o] = sum_{i,k : i*s - p + k = o} in[i] w[k]: for output parity r the valid
    taps have (r + p - k) % s == 0 and read in[q + (r + p - k)//s] for o = q*s + r."""
    phases, perm = [], []
    for ry in range(sy):
        for rx in range(sx):
            taps, dys, dxs = [], [], []
            for ky in range(kh):
                if (ry + py - ky) % sy:
                    continue
                for kx in range(kw):
                    if (rx + px - kx) % sx:
                        continue
                    taps.append(ky * kw + kx)
                    dys.append((ry + py - ky) // sy)
                    dxs.append((rx + px - kx) // sx)
            phases.append(_Phase(taps, dys, dxs, 1, 1, sy, sx, ry, rx, len(perm)))
            perm += taps
    return phases, perm


def _direct_phase(kh, kw, sy, sx, py, px):
    taps = list(range(kh * kw))
    dys = [t // kw - py for t in taps]
    dxs = [t % kw - px for t in taps]
    return [_Phase(taps, dys, dxs, sy, sx, 1, 1, 0, 0, 0)], taps


class GemmLayer:
    """One contraction weight of the model.

    kind "conv"   : nn.Conv2d weight [Cout, Cin, kh, kw]   (graph/encodingBlock.py, graph/decoder.py:79,122,172)
    kind "convT"  : nn.ConvTranspose2d weight [Cin, Cout, kh, kw] (graph/decoder.py:12-15,43-46,73-77,116-120)
    kind "linear" : nn.Linear weight [out, in] == 1x1 conv on a 1x1 map (graph/encoder.py:22, decoder.py:166-167)
    """

    def __init__(self, kind: str, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter] = None,
                 kernel=(1, 1), stride=(1, 1), padding=(0, 0), output_padding=(0, 0)):
        self.kind, self.weight, self.bias = kind, weight, bias
        self.kh, self.kw = kernel
        self.sy, self.sx = stride
        self.py, self.px = padding
        self.opy, self.opx = output_padding
        if kind == "convT":
            self.Cin, self.Cout = weight.shape[0], weight.shape[1]
        else:
            self.Cout, self.Cin = weight.shape[0], weight.shape[1]
        self.T = self.kh * self.kw
        T, Cin, Cout = self.T, self.Cin, self.Cout
        if kind == "convT":
            self.f_phases, self.f_perm = _transposed_phases(self.kh, self.kw, self.sy, self.sx, self.py, self.px)
            self.d_phases, self.d_perm = _direct_phase(self.kh, self.kw, self.sy, self.sx, self.py, self.px)
            # W[ci][co][t]: forward operand rows = co, K = (tap, ci); dgrad operand rows = ci, K = (tap, co)
            self.f_src = (T, Cout * T)
            self.d_src = (Cout * T, T)
        else:
            self.f_phases, self.f_perm = _direct_phase(self.kh, self.kw, self.sy, self.sx, self.py, self.px)
            self.d_phases, self.d_perm = _transposed_phases(self.kh, self.kw, self.sy, self.sx, self.py, self.px)
            # W[co][ci][t]
            self.f_src = (Cin * T, T)
            self.d_src = (T, Cin * T)
        weight._bvae_layer = self   # lets FlatParams find every contraction weight for the batched repack
        self._wf = self._wd = None
        self._wscratch = None       # zero-in / zero-out fp32 scratch for the packed weight-gradient reduction
        self._wf_key = self._wd_key = None
        self._cache: Dict[tuple, object] = {}

    # ---- geometry -------------------------------------------------------------------------------------
    def out_hw(self, H: int, W: int) -> Tuple[int, int]:
        if self.kind == "convT":
            return ((H - 1) * self.sy - 2 * self.py + self.kh + self.opy,
                    (W - 1) * self.sx - 2 * self.px + self.kw + self.opx)
        return ((H + 2 * self.py - self.kh) // self.sy + 1, (W + 2 * self.px - self.kw) // self.sx + 1)

    # ---- packed operands ------------------------------------------------------------------------------
    def _key(self):
        """identity of the fp32 master the packed operands were made from: the global epoch (anything that rewrote
        parameters behind autograd's back: load, broadcast) and the epoch of the weight's own flat bucket (its fused Adam
        step) -- per bucket, so that stepping ONE module (a discriminator, another model replica) does not make every other
        module re-pack its operands"""
        w = self.weight
        flat = getattr(w, "_bvae_flat", None)
        return (w.data_ptr(), w._version, _PARAM_EPOCH[0], flat.epoch if flat is not None else 0)

    def _pack(self, rows: int, cc: int, src, perm) -> torch.Tensor:
        T = len(perm)
        dst = torch.empty((rows, T * cc), dtype=BF16, device=self.weight.device)
        arr = (C.c_int32 * T)(*perm)
        w = self.weight.detach()
        assert w.is_contiguous() and w.dtype == torch.float32
        _lib.check(_lib.lib().bvae_pack_weight(w.data_ptr(), dst.data_ptr(), rows, T, cc, src[0], src[1], arr,
                                               T * cc, _lib.stream_ptr()), "pack_weight")
        return dst

    def w_fwd(self) -> torch.Tensor:
        k = self._key()
        if self._wf_key != k:
            self._wf = self._pack(self.Cout, self.Cin, self.f_src, self.f_perm)
            self._wf_key = k
        return self._wf

    def w_dgrad(self) -> torch.Tensor:
        k = self._key()
        if self._wd_key != k:
            self._wd = self._pack(self.Cin, self.Cout, self.d_src, self.d_perm)
            self._wd_key = k
        return self._wd

    def channel_class(self) -> str:
        """contraction launches by their narrower channel count (which decides what bounds them)"""
        c = min(self.Cin, self.Cout)
        return "ch>=256" if c >= 256 else ("ch128" if c >= 128 else "ch<=64")

    def algorithmic_flops(self, n: int, hw_in: int, hw_out: int) -> float:
        """SURVEY.md section 8 convention, one pass (forward == data gradient == weight gradient):
        Conv 2*Cout*Hout*Wout*Cin*kh*kw, ConvT 2*Cin*Hin*Win*Cout*kh*kw, Linear 2*in*out -- per sample, times n.
        hw_in / hw_out: pixels of the layer's input / output map."""
        hw = hw_in if self.kind == "convT" else hw_out
        return 2.0 * n * hw * self.Cin * self.Cout * self.kh * self.kw

    # ---- launches -------------------------------------------------------------------------------------
    def _run_phases(self, phases, wpk: torch.Tensor, x: Act, y: Act, cout: int, grid_of, bias, act, slope,
                    addend: Optional[Act], mask: Optional[Act], mask_slope: float, tag: str, want_stats: bool = False):
        """returns the fused-statistics buffer (zeroed, then filled by the epilogues of all phases) or None"""
        lib = _lib.lib()
        st = _lib.stream_ptr()
        key = (tag, x.N, x.H, x.W, x.pitch, y.H, y.W, y.pitch, y.f32, act, slope, addend is not None and addend.pitch,
               mask is not None and mask.pitch, mask_slope, bias is not None)
        stats = None
        if want_stats:
            ok = self._cache.get(("stats_ok",) + key)
            if ok is None or ok:
                stats = torch.zeros(y.N * cout * 6, dtype=torch.float32, device=y.t.device)
        descs = self._cache.get(key)
        if descs is None:
            descs = []
            live = []
            for ph in phases:
                if not ph.taps:
                    raise RuntimeError("phase without taps (kernel smaller than stride) is not supported")
                QH, QW = grid_of(ph)
                if QH > 0 and QW > 0:
                    live.append((ph, QH, QW))

            def base_desc(ph):
                d = ConvDesc()
                d.N, d.H, d.W, d.C, d.x_pitch = x.N, x.H, x.W, x.C, x.pitch
                d.Cout, d.w_pitch = cout, wpk.shape[1]
                d.sy, d.sx = ph.sy, ph.sx
                d.OH, d.OW, d.y_pitch = y.H, y.W, y.pitch
                d.osy, d.osx = ph.osy, ph.osx
                d.add_pitch = addend.pitch if addend is not None else 0
                d.mask_pitch = mask.pitch if mask is not None else 0
                d.act, d.out_f32, d.slope, d.mask_slope = int(act), int(y.f32), float(slope), float(mask_slope)
                return d

            ntot = sum(len(ph.taps) for ph, _, _ in live)
            contiguous = all(live[i][0].k_off + len(live[i][0].taps) == live[i + 1][0].k_off for i in range(len(live) - 1))
            if len(live) > 1 and len(live) <= _lib.MAX_PHASES and ntot <= _lib.MAX_TAPS and contiguous:
                # all sub-pixel phases in ONE launch: tiles of the same input patch run back to back (L2 reuse)
                d = base_desc(live[0][0])
                d.nphase, d.ntaps = len(live), ntot
                t = 0
                for i, (ph, QH, QW) in enumerate(live):
                    d.ph_ntaps[i], d.ph_ooy[i], d.ph_oox[i], d.ph_QH[i], d.ph_QW[i] = len(ph.taps), ph.ooy, ph.oox, QH, QW
                    for j in range(len(ph.taps)):
                        d.dy[t], d.dx[t] = ph.dy[j], ph.dx[j]
                        t += 1
                d.QH, d.QW = max(q[1] for q in live), max(q[2] for q in live)
                descs.append((d, live[0][0].k_off * x.C * 2))
            else:
                for ph, QH, QW in live:
                    d = base_desc(ph)
                    d.ntaps = len(ph.taps)
                    for i in range(len(ph.taps)):
                        d.dy[i], d.dx[i] = ph.dy[i], ph.dx[i]
                    d.QH, d.QW = QH, QW
                    d.ooy, d.oox = ph.ooy, ph.oox
                    descs.append((d, ph.k_off * x.C * 2))
            self._cache[key] = descs
        if stats is not None and ("stats_ok",) + key not in self._cache:
            # every phase must be able to fuse (tcgen05 path, fp32 output, each M tile inside one sample)
            ok = _IMPL[0] != _lib.IMPL_SIMT
            for d, woff in descs:
                d.x, d.w, d.y = x.ptr, wpk.data_ptr() + woff, y.ptr
                ok = ok and bool(lib.bvae_conv_stats_ok(C.byref(d)))
            self._cache[("stats_ok",) + key] = ok
            if not ok:
                stats = None
        xp, yp, wp = x.ptr, y.ptr, wpk.data_ptr()
        sp = stats.data_ptr() if stats is not None else None
        bp = bias.data_ptr() if bias is not None else None
        ap = addend.ptr if addend is not None else None
        mp = mask.ptr if mask is not None else None
        fl, cl = 0.0, None
        if profiling():
            # forward: x is the layer's input, y its output; data gradient: x is dy (output side), y is dx (input side)
            hw_in, hw_out = (x.H * x.W, y.H * y.W) if tag == "f" else (y.H * y.W, x.H * x.W)
            fl, cl = self.algorithmic_flops(x.N, hw_in, hw_out) / max(1, len(descs)), self.channel_class()
            if self.Cin == 1 and tag == "d":
                fl = 0.0                     # the data gradient of a C_in = 1 stem is never needed
        for d, woff in descs:
            d.x, d.w, d.y, d.bias, d.addend, d.mask, d.stats = xp, wp + woff, yp, bp, ap, mp, sp
            _lib.check(_timed("conv_gemm", lib.bvae_conv_gemm, C.byref(d), _IMPL[0], st,
                              detail="%s %s %d->%d k%dx%d s%dx%d in%dx%d" % (tag, self.kind, self.Cin, self.Cout, self.kh,
                                                                             self.kw, self.sy, self.sx, x.H, x.W),
                              flops=fl, cls=cl),
                       "conv_gemm[%s]" % tag)
        return stats

    def forward(self, x: Act, y: Act, act: bool = False, slope: float = 0.0, use_bias: bool = True,
                want_stats: bool = False):
        """y = epi(conv(x)); y geometry must be out_hw(x) with C == Cout.  With want_stats the InstanceNorm statistics
        of y are accumulated by the epilogue when possible; returns that buffer (for NormBlock.forward) or None."""
        assert x.C == self.Cin and y.C == self.Cout, (x.C, self.Cin, y.C, self.Cout)
        if self.kind == "convT":
            grid_of = lambda ph: (-(-(y.H - ph.ooy) // ph.osy), -(-(y.W - ph.oox) // ph.osx))
        else:
            grid_of = lambda ph: (y.H, y.W)
        bias = self.bias.detach() if (self.bias is not None and use_bias) else None
        fuse = want_stats and _FUSE_STATS[0] and y.f32 and not act and y.H * y.W > 128
        return self._run_phases(self.f_phases, self.w_fwd(), x, y, self.Cout, grid_of, bias, act, slope, None, None, 0.0,
                                "f", fuse)

    def dgrad(self, dy: Act, dx: Act, addend: Optional[Act] = None, mask: Optional[Act] = None,
              mask_slope: float = 0.0):
        """dx = conv_backward_data(dy) (+ addend) (* act'(mask))."""
        assert dy.C == self.Cout and dx.C == self.Cin
        if self.kind == "convT":
            grid_of = lambda ph: (dx.H, dx.W)
        else:
            grid_of = lambda ph: (-(-(dx.H - ph.ooy) // ph.osy), -(-(dx.W - ph.oox) // ph.osx))
        self._run_phases(self.d_phases, self.w_dgrad(), dy, dx, self.Cin, grid_of, None, False, 0.0, addend, mask,
                         mask_slope, "d")

    def wgrad(self, x: Act, dy: Act):
        """weight.grad += conv_backward_weight(x, dy)  (on the companion stream inside engine.async_wgrad())."""
        if _WGRAD_ASYNC[0] and self.weight.requires_grad:
            cur, ws = _wgrad_pair()
            ws.wait_stream(cur)                      # x and dy are complete at this point of the current stream
            _WGRAD_KEEP.setdefault((cur.device.index, cur.cuda_stream), []).append((x.t, dy.t))
            with torch.cuda.stream(ws):
                self._wgrad_launch(x, dy)
            return
        self._wgrad_launch(x, dy)

    def _wgrad_launch(self, x: Act, dy: Act):
        lib = _lib.lib()
        st = _lib.stream_ptr()
        if self.weight.requires_grad:
            a, s = (x, dy) if self.kind == "convT" else (dy, x)
            key = ("w", a.N, a.H, a.W, a.pitch, s.H, s.W, s.pitch)
            d = self._cache.get(key)
            if d is None:
                d = WgradDesc()
                d.N, d.AH, d.AW, d.Ca, d.a_pitch = a.N, a.H, a.W, a.C, a.pitch
                d.SH, d.SW, d.Cs, d.s_pitch = s.H, s.W, s.C, s.pitch
                d.sy, d.sx, d.ntaps, d.T = self.sy, self.sx, self.T, self.T
                for t in range(self.T):
                    d.dy[t], d.dx[t], d.tap_idx[t] = t // self.kw - self.py, t % self.kw - self.px, t
                self._cache[key] = d
            d.a, d.s, d.dw = a.ptr, s.ptr, grad_ptr(self.weight)
            if self.T > 1 and self.Cin % 32 == 0 and self.Cout % 32 == 0:
                if self._wscratch is None or self._wscratch.device != self.weight.device:
                    self._wscratch = torch.zeros(self.weight.numel(), dtype=torch.float32, device=self.weight.device)
                d.scratch = self._wscratch.data_ptr()
            fl, cl = 0.0, None
            if profiling():
                fl, cl = self.algorithmic_flops(x.N, x.H * x.W, dy.H * dy.W), self.channel_class()
            _lib.check(_timed("wgrad_gemm", lib.bvae_wgrad_gemm, C.byref(d), _IMPL[0], st,
                              detail="w %s %d->%d k%dx%d s%dx%d in%dx%d" % (self.kind, self.Cin, self.Cout, self.kh, self.kw,
                                                                            self.sy, self.sx, x.H, x.W),
                              flops=fl, cls=cl), "wgrad_gemm")

    def bias_grad(self, dy: Act):
        if self.bias is not None and self.bias.requires_grad:
            _lib.check(_lib.lib().bvae_colsum(dy.ptr, int(dy.f32), dy.N * dy.H * dy.W, dy.C, dy.pitch,
                                              grad_ptr(self.bias), _lib.stream_ptr()), "colsum")

    def zero_bias_grad(self):
        """Bias in front of an InstanceNorm: its gradient is analytically zero (the reference only yields rounding
        noise there); make sure a zero .grad exists so optimisers see the same parameter set."""
        if self.bias is not None and self.bias.requires_grad:
            grad_ptr(self.bias)


# ---------------------------------------------------------------------------------------------------------
# InstanceNorm (+CBAM) (+residual) + activation
# ---------------------------------------------------------------------------------------------------------
class NormBlock:
    def __init__(self, C: int, gamma, beta, cbam=None, res_mode: int = 0, slope: float = 0.0):
        """cbam = (w1 [C/16,C,1,1], w2 [C,C/16,1,1], wsp [1,2,3,3]) parameters or None."""
        self.C, self.gamma, self.beta, self.cbam, self.res_mode, self.slope = C, gamma, beta, cbam, res_mode, slope

    def _desc(self, y_pitch, y_f32, out: Act, N, H, W) -> NbDesc:
        d = NbDesc()
        d.N, d.H, d.W, d.C = N, H, W, self.C
        d.y_pitch, d.out_pitch = y_pitch, out.pitch
        d.has_cbam = int(self.cbam is not None)
        d.Cr = self.C // 16
        d.res_mode, d.y_f32, d.slope, d.eps = self.res_mode, int(y_f32), self.slope, 1e-5
        d.gamma, d.beta = self.gamma.data_ptr(), self.beta.data_ptr()
        if self.cbam is not None:
            d.w1, d.w2, d.wsp = (p.data_ptr() for p in self.cbam)
        return d

    def forward(self, y: Act, out: Act, res: Optional[Act] = None, stats: Optional[torch.Tensor] = None) -> dict:
        N, H, W, Cc = y.N, y.H, y.W, self.C
        dev = y.t.device
        ctx = {"N": N, "H": H, "W": W, "out": out,
               "uhat": torch.empty((N, H, W, Cc), dtype=BF16, device=dev) if _SAVE_FOR_BACKWARD else None,
               "nc": torch.empty((N, Cc, 8), dtype=torch.float32, device=dev),
               "nc_idx": torch.empty((N, Cc), dtype=torch.int32, device=dev)}
        fused = stats is not None
        if stats is None:
            stats = torch.empty((N * Cc * 6,), dtype=torch.float32, device=dev)
        d = self._desc(y.pitch, y.f32, out, N, H, W)
        d.stats_fused = int(fused)
        d.y, d.out, d.stats = y.ptr, out.ptr, stats.data_ptr()
        d.uhat = ctx["uhat"].data_ptr() if ctx["uhat"] is not None else None      # NULL: inference, not written
        d.nc, d.nc_idx = ctx["nc"].data_ptr(), ctx["nc_idx"].data_ptr()
        if self.cbam is not None:
            ctx["sa"] = torch.empty((N, H * W, 2), dtype=torch.float32, device=dev)
            ctx["cidx"] = torch.empty((N, H * W), dtype=torch.int32, device=dev)
            ctx["gs"] = torch.empty((N, H * W), dtype=torch.float32, device=dev)
            d.sa, d.cidx, d.gs = ctx["sa"].data_ptr(), ctx["cidx"].data_ptr(), ctx["gs"].data_ptr()
        if self.res_mode == 2:
            assert res is not None
            d.res, d.res_pitch = res.ptr, res.pitch
        _lib.check(_timed("nb_forward", _lib.lib().bvae_nb_forward, C.byref(d), _lib.stream_ptr(),
                          detail="C%d %dx%d cbam%d" % (Cc, H, W, int(self.cbam is not None))), "nb_forward")
        return ctx

    def backward(self, ctx: dict, dout: Act, dy: Act, dres: Optional[Act] = None):
        N, H, W, Cc = ctx["N"], ctx["H"], ctx["W"], self.C
        dev = dout.t.device
        out: Act = ctx["out"]
        d = self._desc(0, True, out, N, H, W)
        if ctx["uhat"] is None:
            raise RuntimeError("norm block: backward of a forward pass that ran under engine.saving(False)")
        d.uhat, d.out, d.nc, d.nc_idx = ctx["uhat"].data_ptr(), out.ptr, ctx["nc"].data_ptr(), ctx["nc_idx"].data_ptr()
        bwd_nc = torch.empty((N, Cc, 4), dtype=torch.float32, device=dev)
        d.bwd_nc = bwd_nc.data_ptr()
        d.dout, d.dout_pitch, d.dy, d.dy_pitch = dout.ptr, dout.pitch, dy.ptr, dy.pitch
        d.dgamma, d.dbeta = grad_ptr(self.gamma), grad_ptr(self.beta)
        if self.cbam is not None:
            bwd_px = torch.empty((N, H * W, 4), dtype=torch.float32, device=dev)
            d.bwd_px = bwd_px.data_ptr()
            bwd_h = torch.empty((N, 192), dtype=torch.float32, device=dev)
            d.bwd_h = bwd_h.data_ptr()
            d.sa, d.cidx, d.gs = ctx["sa"].data_ptr(), ctx["cidx"].data_ptr(), ctx["gs"].data_ptr()
            d.dw1, d.dw2, d.dwsp = (grad_ptr(p) for p in self.cbam)
        if self.res_mode == 2 and dres is not None:
            d.dres, d.dres_pitch = dres.ptr, dres.pitch
        _lib.check(_timed("nb_backward", _lib.lib().bvae_nb_backward, C.byref(d), _lib.stream_ptr(),
                          detail="C%d %dx%d cbam%d" % (Cc, H, W, int(self.cbam is not None))), "nb_backward")


# ---------------------------------------------------------------------------------------------------------
# flat parameter / gradient buckets
# ---------------------------------------------------------------------------------------------------------
class FlatParams:
    """All parameters of a module in ONE fp32 buffer (and their gradients in another), each parameter a view.

    Replaces the per-parameter Adam loop (agent/barGen.py:61-62,327,333: 221 tensors) by one launch, and the
    per-parameter Horovod hooks (agent/barGen_horovod.py:91-99) by bucketed all-reduces of contiguous slices."""
    ALIGN = 64

    def __init__(self, module: torch.nn.Module):
        params, seen = [], set()
        for p in module.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.params = params
        dev = params[0].device
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.numel = off
        self.epoch = 0             # bumped by every fused Adam step on this bucket (GemmLayer._key)
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = self.exp_avg_sq = None
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                v = self.data[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                p._bvae_flat = self
                p._bvae_off = o
                p.grad = None
        bump_param_epoch()

    def grad_view(self, p) -> torch.Tensor:
        o = p._bvae_off
        return self.grad[o:o + p.numel()].view(p.shape)

    def attach_grads(self, zero: bool = True):
        """Point every .grad at its slice of the flat gradient bucket (one memset instead of 221).  The view objects
        are built once and re-attached only where ``p.grad`` is no longer that object: creating 221 views every step
        cost ~1 ms of host time in front of the step's first kernel -- GPU idle time in a loop that reads the loss back
        every step (agent/barGen.py:335)."""
        if zero:
            self.grad.zero_()
        views = getattr(self, "_grad_views", None)
        if views is None:
            views = self._grad_views = [self.grad_view(p) for p in self.params]
        for p, v in zip(self.params, views):
            if p.requires_grad and p.grad is not v:
                p.grad = v

    def intact(self) -> bool:
        p, o = self.params[0], self.offsets[0]
        return p.data_ptr() == self.data.data_ptr() + o * 4


def flatten(module: torch.nn.Module) -> FlatParams:
    flat = getattr(module, "_bvae_flat_params", None)
    if flat is None or not flat.intact() or flat.data.device != next(module.parameters()).device:
        flat = FlatParams(module)
        module._bvae_flat_params = flat
    return flat


class PackPlan:
    """One-launch repack of every bf16 contraction operand that the model has used so far (bvae_pack_plan_*).

    Built after the first optimiser step from the operands the layers packed lazily during that step; it takes over
    those buffers and rewrites them in place after every later step."""

    def __init__(self, flat: FlatParams):
        self.handle = None
        self.items = []          # (layer, which, tensor, weight data_ptr)
        jobs = []
        for p in flat.params:
            layer = getattr(p, "_bvae_layer", None)
            if layer is None or layer.weight is not p:
                continue
            for which, buf, rows, cc, src, perm in (("f", layer._wf, layer.Cout, layer.Cin, layer.f_src, layer.f_perm),
                                                    ("d", layer._wd, layer.Cin, layer.Cout, layer.d_src, layer.d_perm)):
                if buf is None:
                    continue
                j = _lib.PackJob()
                j.src, j.dst = p.data_ptr(), buf.data_ptr()
                j.R, j.T, j.Cc, j.dst_pitch = rows, len(perm), cc, buf.shape[1]
                j.sr, j.sc = src
                for i, t in enumerate(perm):
                    j.perm[i] = t
                jobs.append(j)
                self.items.append((layer, which, buf, p.data_ptr()))
        if jobs:
            arr = (_lib.PackJob * len(jobs))(*jobs)
            h = _lib.c_vp()
            _lib.check(_lib.lib().bvae_pack_plan_create(arr, len(jobs), C.byref(h)), "pack_plan_create")
            self.handle = h

    def valid(self) -> bool:
        for layer, which, buf, ptr in self.items:
            if layer.weight.data_ptr() != ptr or (layer._wf if which == "f" else layer._wd) is not buf:
                return False
        return True

    def run(self):
        if self.handle is None:
            return
        _lib.check(_lib.lib().bvae_pack_plan_run(self.handle, _lib.stream_ptr()), "pack_plan_run")
        for layer, which, _, _ in self.items:
            if which == "f":
                layer._wf_key = layer._key()
            else:
                layer._wd_key = layer._key()

    def __del__(self):
        if self.handle is not None:
            try:
                _lib.load().bvae_pack_plan_destroy(self.handle)
            except Exception:
                pass
            self.handle = None


def repack_weights(flat: FlatParams):
    """Refresh the bf16 operands of all contraction weights in one launch (after anything rewrote flat.data)."""
    plan = getattr(flat, "pack_plan", None)
    if plan is None or not plan.valid():
        plan = flat.pack_plan = PackPlan(flat)
    plan.run()


def adam_step(flat: FlatParams, lr: float, step: int, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0,
              repack: bool = True):
    """torch.optim.Adam semantics (agent/barGen.py:61-62) on the flat bucket, one kernel (+ one repack launch)."""
    if flat.exp_avg is None:
        flat.exp_avg = torch.zeros_like(flat.data)
        flat.exp_avg_sq = torch.zeros_like(flat.data)
    _lib.check(_lib.lib().bvae_adam_step(flat.data.data_ptr(), flat.grad.data_ptr(), flat.exp_avg.data_ptr(),
                                         flat.exp_avg_sq.data_ptr(), flat.numel, lr, betas[0], betas[1], eps, step,
                                         grad_scale, _lib.stream_ptr()), "adam")
    flat.epoch += 1
    if repack:
        repack_weights(flat)


def adam_step_dev(flat: FlatParams, hyper_dev: torch.Tensor, repack: bool = True):
    """adam_step with the step-dependent scalars read from device memory (bvae_adam_step_dev): the launch is identical
    every step, so it can live inside a captured CUDA graph; ``adam_hyper_upload`` sets the scalars before each replay."""
    if flat.exp_avg is None:
        flat.exp_avg = torch.zeros_like(flat.data)
        flat.exp_avg_sq = torch.zeros_like(flat.data)
    _lib.check(_lib.lib().bvae_adam_step_dev(flat.data.data_ptr(), flat.grad.data_ptr(), flat.exp_avg.data_ptr(),
                                             flat.exp_avg_sq.data_ptr(), flat.numel, hyper_dev.data_ptr(),
                                             _lib.stream_ptr()), "adam_dev")
    flat.epoch += 1
    if repack:
        repack_weights(flat)


def adam_hyper_upload(hyper_dev: torch.Tensor, lr: float, step: int, betas=(0.9, 0.999), eps: float = 1e-8,
                      grad_scale: float = 1.0):
    _lib.check(_lib.lib().bvae_adam_hyper_upload(lr, betas[0], betas[1], eps, step, grad_scale, hyper_dev.data_ptr(),
                                                 _lib.stream_ptr()), "adam_hyper_upload")


def capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
