"""Execution engine of the score network: NHWC bf16 activations, CUDA kernels only.

The reference evaluates NCSNpp.forward (models/ncsnpp.py:238-388) as ~830 eager ATen ops on NCHW
fp32 tensors. Here the same dataflow is *planned once* per (network, batch shape): weights are
packed into the K-major bf16 layout of the tcgen05 conv/GEMM kernel, activations live in a pool of
NHWC bf16 buffers, and the forward pass becomes a fixed launch list of libcsd_b200 kernels
(~O(10) per block) that can be captured into one CUDA graph.

torch is used for memory and streams only; every arithmetic op is a libcsd_b200 kernel and there is
no fallback: a missing library or an unsupported configuration raises.
"""
import math
import os

import torch

from . import kernels as K
from ._lib import CsdError

BF16 = torch.bfloat16
SQRT1_2 = 1.0 / math.sqrt(2.0)


class Act:
    """NHWC bf16 activation: tensor [B, H, W, pitch] with `c` valid channels and, once some GroupNorm
    needs them, its per-channel (sum, sum of squares) [B, c, 2]."""

    __slots__ = ("t", "c", "sums", "pending")

    def __init__(self, t, c, sums=None, pending=None):
        # pending = (partials [B * tiles, c, 2], tiles per image, sums slot): the transposed convolution that wrote the
        # tensor left per-tile partial sums; whoever needs the channel sums first reduces them (BlockOps.ensure_sums /
        # gn_coeffs) - folded into the coefficient kernel when that consumer is a fused GroupNorm prologue
        self.t, self.c, self.sums, self.pending = t, c, sums, pending

    @property
    def shape(self):
        return self.t.shape

    @property
    def pitch(self):
        return self.t.shape[-1]


class BufferPool:
    """Stream-ordered reuse of activation buffers (all work is issued on one stream in plan order)."""

    def __init__(self, device, dtype=BF16):
        self.device = device
        self.dtype = dtype          # activation storage of the plan: bf16, or fp32 for the tf32 plan
        self.free = {}
        self.all = []

    def get(self, shape, dtype=None):
        dtype = dtype or self.dtype
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        t = torch.empty(shape, device=self.device, dtype=dtype)
        self.all.append(t)
        return t

    def put(self, t):
        self.free.setdefault((tuple(t.shape), t.dtype), []).append(t)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.all)


def _groups(c):
    return min(c // 4, 32)


class WSrc:
    """Where one block of a packed weight comes from: a live parameter (so the pack can be refreshed in place after
    an optimizer step) and where its gradient goes. kind: 'conv' = nn.Conv2d weight [Cout, Cin, kh, kw] (optionally a
    slice [ci_off, ci_off + ci_cnt) of its input channels), 'nin' = NIN.W [in, out] (models/layers.py:555-564) used as
    a 1x1 conv, 'eye' = identity (a residual carried as a K segment; no parameter), 'tap' = tap `tap` of a 3x3
    nn.Conv2d weight as a 1x1 convolution [Cout, Cin, 1, 1] (tap-stacked output heads)."""

    __slots__ = ("param", "kind", "ci_off", "ci_cnt", "tap")

    def __init__(self, param, kind="conv", ci_off=0, ci_cnt=None, tap=0):
        self.param, self.kind, self.ci_off, self.ci_cnt, self.tap = param, kind, ci_off, ci_cnt, tap

    def weight(self, device):
        """[Cout, Cin_slice, kh, kw] view of the live parameter."""
        if self.kind == "eye":
            c = self.param
            return torch.eye(c, device=device, dtype=torch.float32).view(c, c, 1, 1)
        w = self.param.detach()
        if self.kind == "nin":
            w = w.t().reshape(w.shape[1], w.shape[0], 1, 1)
        elif self.kind == "tap":
            kw = w.shape[3]
            w = w[:, :, self.tap // kw, self.tap % kw].reshape(w.shape[0], w.shape[1], 1, 1)
        if self.ci_cnt is not None:
            w = w[:, self.ci_off:self.ci_off + self.ci_cnt]
        return w


def _as_segments(weights):
    """Accept the older call form (a list of weight tensors / parameters) as well as lists of WSrc stacks."""
    segs = []
    for w in weights:
        if isinstance(w, WSrc):
            segs.append([w])
        elif isinstance(w, (list, tuple)):
            segs.append(list(w))
        else:
            segs.append([WSrc(w)])
    return segs


class PackedConv:
    """bf16 K-major weights [n_pad, k_total] + fp32 bias [n_pad + 16] for csd_conv_gemm.

    segments: one entry per K segment; an entry is a WSrc or a list of WSrc stacked along the output channels (the
    attention q|k projection). bias: a parameter / tensor, or a list of (parameter, co_off) that are summed into the
    bias vector (Conv_1.bias + Conv_2.bias when the skip convolution rides as a K segment). The device buffers are
    persistent: refresh() re-packs them in place from the live parameters, so recorded launch lists stay valid
    across optimizer steps."""

    def __init__(self, weights, bias, device, dtype=BF16):
        self.segs = _as_segments(weights)
        self.device = device
        self.dtype = dtype          # operand storage: bf16 (kind::f16) or fp32 (kind::tf32)
        w0 = [s.weight(device) for s in self.segs[0]]
        cout = sum(w.shape[0] for w in w0)
        self.cout = cout
        self.n_store = K.ceil_to(cout, 8)
        n16 = K.ceil_to(cout, 16)
        if n16 <= 256:
            self.n_tile = n16
        else:
            self.n_tile = K.ceil_to((n16 + 1) // 2, 16)
            if self.n_tile * 2 > 512:
                self.n_tile = 256
        n_tiles = math.ceil(self.n_store / self.n_tile)
        self.n_pad = n_tiles * self.n_tile
        if bias is None:
            self.bias_srcs = []
        elif isinstance(bias, (list, tuple)):
            self.bias_srcs = [(b, off) for b, off in bias]
        else:
            self.bias_srcs = [(bias, 0)]
        self.wt = None
        self.bias = torch.zeros(self.n_pad + 16, device=device, dtype=torch.float32)
        self.identity_skip = False
        self._dgrads = {}
        PackedConv.generation += 1
        self.refresh()

    generation = 0     # bumped whenever a packed operand is created: invalidates the engine's batched pack table

    @staticmethod
    def _block_jobs(src, dst, dst_off, k_pad, dst_pitch, dgrad, scale):
        """csd_pack_weights job of one WSrc block, or None when the block cannot take the fast path."""
        p = src.param
        if src.kind == "eye" or not (torch.is_tensor(p) and p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            return None
        if src.kind == "nin":
            cin_tot, cout = p.shape
            taps, cin = 1, (src.ci_cnt if src.ci_cnt is not None else cin_tot)
            s_co, s_ci, s_tap, off = 1, cout, 0, src.ci_off * cout
        elif src.kind == "tap":
            cout, cin_tot = p.shape[0], p.shape[1]
            ntap = p.shape[2] * p.shape[3]
            taps, cin = 1, (src.ci_cnt if src.ci_cnt is not None else cin_tot)
            s_co, s_ci, s_tap, off = cin_tot * ntap, ntap, 0, src.ci_off * ntap + src.tap
        else:
            cout, cin_tot = p.shape[0], p.shape[1]
            taps = p.shape[2] * p.shape[3]
            cin = src.ci_cnt if src.ci_cnt is not None else cin_tot
            s_co, s_ci, s_tap, off = cin_tot * taps, taps, 1, src.ci_off * taps
        if dgrad:
            return K.pack_job(p.detach(), dst, cin, cout, taps, k_pad, dst_pitch, s_ci, s_co, s_tap, flip=True, scale=scale,
                              src_off=off, dst_off=dst_off), cout
        return K.pack_job(p.detach(), dst, cout, cin, taps, k_pad, dst_pitch, s_co, s_ci, s_tap, src_off=off,
                          dst_off=dst_off), cout

    def jobs(self):
        """Job list that refreshes this operand (weights, bias, data-gradient packs), or None -> torch refresh."""
        out = []
        k_total = self.wt.shape[1]
        koff = 0
        for i, seg in enumerate(self.segs):
            w0 = seg[0].weight(self.device)
            taps, cin = w0.shape[2] * w0.shape[3], w0.shape[1]
            cpad = K.ceil_to(cin, 32)
            co = 0
            for src in seg:
                if src.kind == "eye":
                    co += src.param
                    continue
                r = self._block_jobs(src, self.wt, co * k_total + koff, cpad, k_total, False, 1.0)
                if r is None:
                    return None
                out.append(r[0])
                co += r[1]
            koff += taps * cpad
        by_off = {}
        for b, off in self.bias_srcs:
            if not (b.is_cuda and b.dtype == torch.float32 and b.is_contiguous()):
                return None
            by_off.setdefault(off, []).append(b.detach())
        for off, bs in by_off.items():
            if len(bs) > 2 or (len(bs) == 2 and bs[0].numel() != bs[1].numel()):
                return None
            out.append(K.bias_job(self.bias, off, bs[0], bs[1] if len(bs) == 2 else None))
        for d in self._dgrads.values():
            seg = self.segs[d.i]
            cout_total = sum(s.weight(self.device).shape[0] for s in seg)
            k_pad = K.ceil_to(cout_total, 32)
            w0 = seg[0].weight(self.device)
            taps = w0.shape[2] * w0.shape[3]
            co = 0
            for src in seg:
                r = self._block_jobs(src, d.wt, co, k_pad, taps * k_pad, True, d.scale)
                if r is None:
                    return None
                out.append(r[0])
                co += r[1]
        return out

    def seg_weight(self, i):
        ws = [s.weight(self.device).to(self.device) for s in self.segs[i]]
        return ws[0] if len(ws) == 1 else torch.cat(ws, 0)

    def refresh(self):
        parts = [K.pack_conv_weight(self.seg_weight(i), n_pad=self.n_pad, dtype=self.dtype) for i in range(len(self.segs))]
        wt = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
        if self.wt is None:
            self.wt = wt.contiguous()
        else:
            self.wt.copy_(wt)
        if self.bias_srcs:
            self.bias.zero_()
            for b, off in self.bias_srcs:
                self.bias[off:off + b.shape[0]] += b.detach().to(device=self.device, dtype=torch.float32)
        for d in self._dgrads.values():
            d.refresh()

    def dgrad(self, i, scale=1.0):
        """Packed weights of the data gradient of segment i: [Cin_i, Cout, kh, kw] = scale * W_i flipped in (ky, kx)
        and transposed in (co, ci), so that d a_i = conv(d out, this) with the same 'same' padding."""
        if self.dtype != BF16:
            raise CsdError("the training plan (data-gradient packs) exists for the bf16 plan only")
        key = (i, float(scale))
        if key not in self._dgrads:
            self._dgrads[key] = DgradPack(self, i, scale)
        return self._dgrads[key]


class DgradPack:
    def __init__(self, pc, i, scale):
        self.pc, self.i, self.scale = pc, i, scale
        self.wt = None
        PackedConv.generation += 1
        w = pc.seg_weight(i)
        self.cout = w.shape[1]                     # output channels of the dgrad = input channels of the segment
        self.n_store = K.ceil_to(self.cout, 8)
        n16 = K.ceil_to(self.cout, 16)
        self.n_tile = n16 if n16 <= 256 else K.ceil_to((n16 + 1) // 2, 16)
        if self.n_tile * 2 > 512 and n16 > 256:
            self.n_tile = 256
        self.n_pad = math.ceil(self.n_store / self.n_tile) * self.n_tile
        self.refresh()

    def refresh(self):
        w = self.pc.seg_weight(self.i).to(torch.float32)
        wd = (w.flip(2, 3).transpose(0, 1) * self.scale).contiguous()
        wt = K.pack_conv_weight(wd, n_pad=self.n_pad)
        if self.wt is None:
            self.wt = wt
        else:
            self.wt.copy_(wt)


def nin_as_conv(W):
    """NIN weight [in, out] (models/layers.py:555-564) -> conv weight [out, in, 1, 1]."""
    return W.detach().t().reshape(W.shape[1], W.shape[0], 1, 1)


class Recorder:
    """Launch list: (fn, args, kwargs) triples executed in order."""

    def __init__(self):
        self.ops = []

    def add(self, fn, *args, **kwargs):
        self.ops.append((fn, args, kwargs))

    def run(self, skip=()):
        """skip: launch functions to leave out (their outputs from an earlier run of the same list are still valid)."""
        for fn, args, kwargs in self.ops:
            if fn in skip:
                continue
            fn(*args, **kwargs)

    def __len__(self):
        return len(self.ops)


# A/B switch for the output heads: 0 = GroupNorm pass + per-tap kernel, 1 = transposed kernel with the pyramid as an
# identity K segment, 2 = transposed kernel, pyramid added by its FIR upsampling.
# 3 (default) = tap-stacked: ONE 1x1 convolution to 9 * Cout rows in the transposed kernel (fused GroupNorm prologue, no
# halo, 54 of 128 M rows for the 6-channel heads) + csd_tap_shift_sum (shifted sum of the 9 partial maps, bias, pyramid).
HEAD_MODE = int(os.environ.get("CSD_HEAD_MODE", "3"))
# fp32 (tf32) plan: 3x3 convolutions of the large levels in the transposed kernel's kind::tf32 instance (fused GroupNorm
# prologue, epilogue statistics, fp32 residual in the epilogue); 0 = everything in the per-tap tf32 kernel.
TF32_TRANSPOSED = os.environ.get("CSD_NO_TF32_TRANSPOSED", "0") != "1"


class BlockOps:
    """Builders that append the kernels of one reference layer to a Recorder."""

    # inference: a GroupNorm whose inputs have no channel sums yet runs as one launch when the shape allows it
    # (K.gn_fused). The training subclass keeps the separate statistics: its backward reads them.
    fuse_small_gn = True
    # inference: the few-channel output heads (GroupNorm + SiLU + conv3x3 -> 3/6 channels) run in the transposed
    # kernel with the GroupNorm in its prologue; the training subclass keeps the plain path (its backward needs the
    # normalised tensor and takes the residual from the epilogue).
    fast_heads = True
    # inference: QK^T / softmax / PV / projection of an attention block in one kernel (csrc/attention.cu); the training
    # subclass keeps the separate GEMMs (its backward re-uses their operands).
    fused_attention = True
    # inference: the reduction of a transposed convolution's per-tile statistics waits for the first consumer and is
    # folded into its coefficient kernel (gn_coeffs_partials); the training subclass finalizes at once.
    defer_finalize = os.environ.get("CSD_NO_DEFER_FINALIZE", "0") != "1"

    def __init__(self, device, pool, rec, stats_arena):
        self.device = device
        self.pool = pool
        self.act_dtype = pool.dtype     # bf16 plan: transposed tcgen05 kernel + fused GroupNorm; fp32 (tf32) plan: per-tap
        self.rec = rec
        self.stats = stats_arena  # fp32 [n_slots, ...] zeroed at the start of every forward
        self.stats_used = 0

    # -- helpers -----------------------------------------------------------------------------
    def _stats_slot(self, batch, channels):
        n = batch * channels * 2
        n_al = (n + 63) // 64 * 64
        if self.stats_used + n_al > self.stats.numel():
            raise CsdError("GroupNorm statistics arena too small")
        s = self.stats[self.stats_used:self.stats_used + n].view(batch, channels, 2)
        self.stats_used += n_al
        return s

    def ensure_sums(self, a):
        """Per-channel sums of a tensor: produced by the convolution that wrote it when that ran in the
        transposed mode, otherwise by one statistics pass - in both cases once per tensor."""
        if a.sums is None and a.pending is not None:
            partials, tiles_img, slot = a.pending
            self.rec.add(K.gn_finalize_partials, partials, slot, a.shape[0], tiles_img, a.c)
            self.pool.put(partials)
            a.sums, a.pending = slot, None
        if a.sums is None:
            a.sums = self._stats_slot(a.shape[0], a.c)
            self.rec.add(K.gn_chan_stats, a.t, a.c, a.sums)
        return a.sums

    def group_norm(self, srcs, gamma, beta, silu, groups=None):
        """srcs: list of 1 or 2 Act (channel concatenation). Returns a new Act. Every GroupNorm output of these
        networks feeds a convolution / GEMM operand (directly or through a FIR pass), so the fp32 plan rounds it to
        tf32 on store (kernels.gn_apply round_out)."""
        b, h, w, _ = srcs[0].shape
        c = sum(a.c for a in srcs)
        groups = groups or _groups(c)
        s0 = srcs[0]
        s1 = srcs[1] if len(srcs) > 1 else None
        if (self.fuse_small_gn and any(a.sums is None for a in srcs)
                and K.gn_fused_supported(s0.c, s1.c if s1 else 0, h * w, groups, b, self.act_dtype)):
            # small levels: statistics + apply in ONE launch (nobody delivers these tensors' sums for free)
            out = self.pool.get((b, h, w, c))
            self.rec.add(K.gn_fused, s0.t, s0.c, s1.t if s1 else None, s1.c if s1 else 0, gamma, beta, out, groups,
                         1e-6, silu, True)
            return Act(out, c)
        sums0 = self.ensure_sums(s0)
        sums1 = self.ensure_sums(s1) if s1 is not None else None
        out = self.pool.get((b, h, w, c))
        self.rec.add(K.gn_apply, s0.t, s0.c, sums0, s1.t if s1 else None, s1.c if s1 else 0, sums1, gamma, beta, out,
                     groups, 1e-6, silu, True)
        return Act(out, c)

    def gn_coeffs(self, srcs, gamma, beta, groups=None):
        """(scale, shift) tables of GroupNorm over the concatenation of `srcs`, one [B, c, 2] tensor per source:
        the fused-prologue replacement of group_norm() for convolutions that run in the transposed mode."""
        b, h, w, _ = srcs[0].shape
        c = sum(a.c for a in srcs)
        s0 = srcs[0]
        s1 = srcs[1] if len(srcs) > 1 else None
        coef0 = self.pool.get((b, s0.c, 2), torch.float32)
        coef1 = self.pool.get((b, s1.c, 2), torch.float32) if s1 is not None else None
        if any(a is not None and a.sums is None and a.pending is not None for a in (s0, s1)):
            # a source still carries its convolution's per-tile partials: reduce them inside the coefficient kernel
            args, done = [], []
            for a in (s0, s1):
                if a is None:
                    args.append(None)
                elif a.sums is None and a.pending is not None:
                    partials, tiles_img, slot = a.pending
                    args.append((None, partials, tiles_img, slot, a.c))
                    done.append((a, partials, slot))
                else:
                    args.append((self.ensure_sums(a), None, 0, None, a.c))
            self.rec.add(K.gn_coeffs_partials, args[0], args[1], gamma, beta, coef0, coef1, b, h * w,
                         groups or _groups(c), 1e-6)
            for a, partials, slot in done:
                self.pool.put(partials)
                a.sums, a.pending = slot, None
            return [coef0] + ([coef1] if s1 is not None else [])
        sums0 = self.ensure_sums(s0)
        sums1 = self.ensure_sums(s1) if s1 is not None else None
        self.rec.add(K.gn_coeffs, sums0, s0.c, sums1, s1.c if s1 else 0, gamma, beta, coef0, coef1, h * w,
                     groups or _groups(c), 1e-6)
        return [coef0] + ([coef1] if s1 is not None else [])

    ragged_tiles = True     # inference plans may run w % 8 != 0 levels in the transposed kernel (bf16, K.T_RAGGED: off)

    def _ragged(self):
        return self.ragged_tiles and self.act_dtype == BF16

    def will_transpose(self, h, w, cout):
        """3x3 stride-1 convolutions with >= 32 output channels on images that tile into 32x8-pixel macro tiles
        run in the persistent transposed kernel (output channels on M, 256 pixels on N); the fp32 plan runs its
        kind::tf32 instance (CSD_NO_TF32_TRANSPOSED=1 keeps the per-tap kernel: A/B switch)."""
        if self.act_dtype != BF16 and not TF32_TRANSPOSED:
            return False
        return K.TRANSPOSED_DEFAULT and cout >= 32 and cout % 8 == 0 and K.transposed_shape_ok(h, w, self._ragged())

    def fusable(self, srcs, cout):
        """GroupNorm+SiLU can ride in the convolution's prologue when the 3x3 conv runs in the transposed mode."""
        _, h, w, _ = srcs[0].shape
        if self.act_dtype != BF16 and not TF32_TRANSPOSED:
            return False
        return (K.FUSE_GN_DEFAULT and K.TRANSPOSED_DEFAULT and cout >= 32 and cout % 8 == 0
                and K.transposed_shape_ok(h, w, self._ragged()) and all(a.c % 8 == 0 for a in srcs))

    def conv(self, segs, pc, out_hw=None, temb=None, temb_pitch=0, res=None, scale=1.0, stride=1, pad=1,
             out=None, head=False, out_pitch=None):
        """segs: list of (Act, taps[, coef]); coef = fused GroupNorm+SiLU table of that segment (gn_coeffs).
        pc: PackedConv. head: few-channel output head forced into the transposed kernel (no statistics).
        Returns Act [B, oh, ow, n_store]."""
        a0 = segs[0][0]
        b, ih, iw, _ = a0.shape
        oh, ow = out_hw if out_hw is not None else (ih, iw)
        if out is None:
            out = self.pool.get((b, oh, ow, pc.n_store))
        seg_list = [(sg[0].t, sg[0].pitch, 0, sg[0].c, sg[1], sg[2] if len(sg) > 2 else None, True) for sg in segs]
        use_t = ((head or self.will_transpose(oh, ow, pc.cout))
                 and K.transposed_eligible(seg_list, oh, ow, stride, pad, allow_1tap=head, ragged=self._ragged()))
        if use_t and res is not None and self.act_dtype == BF16:
            raise CsdError("transposed conv takes its residual as an identity K segment (engine planning error)")
        partials = sums = None
        if use_t and not head:
            # GroupNorm statistics of the output for free: per-(tile, pixel half) partial sums from the epilogue
            tiles_img = math.ceil(oh / K.transposed_tile_rows(oh)) * math.ceil(ow / 8) * 2
            partials = self.pool.get((b * tiles_img, pc.n_store, 2), torch.float32)
            sums = self._stats_slot(b, pc.cout)
        ks, ws = 1, None
        if not use_t:
            ks = K.pick_k_splits(seg_list, b, oh, ow, pc.n_store, pc.n_tile)
            if ks > 1:
                ws = self.pool.get((ks, b * oh * ow, -(-pc.n_store // 8) * 8), torch.float32)
        self.rec.add(K.conv_gemm, seg_list, pc.wt, pc.cout, out, batch=b, h=oh, w=ow, n_store=pc.n_store,
                     n_tile=pc.n_tile, bias=pc.bias, temb=temb, temb_pitch=temb_pitch,
                     res=res.t if res is not None else None, res_pitch=res.pitch if res is not None else 0,
                     scale=scale, stride=stride, pad=pad, in_h=ih, in_w=iw, transposed=use_t, stat_partials=partials,
                     out_pitch=out_pitch, k_splits=ks, splitk_ws=ws)
        if ws is not None:
            self.pool.put(ws)
        if partials is not None and self.defer_finalize:
            return Act(out, pc.cout, None, (partials, tiles_img, sums))
        if partials is not None:
            self.rec.add(K.gn_finalize_partials, partials, sums, b, tiles_img, pc.cout)
            self.pool.put(partials)
        return Act(out, pc.cout, sums)

    def head_tap_stacked(self, pc, hcur):
        b, h, w, _ = hcur.shape
        return (self.act_dtype == BF16 and self.fast_heads and HEAD_MODE == 3 and K.TRANSPOSED_DEFAULT and K.FUSE_GN_DEFAULT
                and K.transposed_shape_ok(h, w, self._ragged()) and hcur.c % 8 == 0 and pc.cout <= 8 and len(pc.segs) == 1
                and len(pc.segs[0]) == 1 and pc.segs[0][0].kind == "conv" and pc.segs[0][0].param.shape[2:] == (3, 3))

    def head_in_transposed_kernel(self, pc, hcur):
        b, h, w, _ = hcur.shape
        return (self.act_dtype == BF16 and self.fast_heads and HEAD_MODE in (1, 2) and K.TRANSPOSED_DEFAULT and K.FUSE_GN_DEFAULT
                and K.transposed_shape_ok(h, w, self._ragged()) and hcur.c % 8 == 0 and pc.n_store == 8 and len(pc.segs) == 1)

    def head(self, gn, pc, hcur, extra, key, res=None):
        """Output head: conv3x3(SiLU(GroupNorm(h))) [+ res] -> the few image channels (models/ncsnpp.py:337-352,
        372-381). Where the image tiles into the transposed kernel's macro tiles the head runs there: M = 128
        channel rows of which 8 are stored, but N = 256 pixels per instruction instead of 16 columns at the
        per-instruction floor of the pixel-major kernel (2x faster, measured), GroupNorm+SiLU in the prologue (no
        normalised copy), the residual pyramid as an identity K segment. extra: dict that owns the derived packs
        (visited by the engine's refresh); key: this head's slot in it."""
        gamma, beta, groups = gn
        if self.head_tap_stacked(pc, hcur) and (res is None or res.c == pc.cout):
            pct = extra.get((key, "taps"))
            if pct is None:
                src = pc.segs[0][0]
                pct = extra[(key, "taps")] = PackedConv([[WSrc(src.param, "tap", tap=t) for t in range(9)]], None,
                                                        self.device, pc.dtype)
            (cf,) = self.gn_coeffs([hcur], gamma, beta, groups)
            part = self.conv([(hcur, 1, cf)], pct, head=True)
            self.pool.put(cf)
            b, h, w, _ = hcur.shape
            out = self.pool.get((b, h, w, pc.n_store))
            self.rec.add(K.tap_shift_sum, part.t, pc.cout, pc.bias, res.t if res is not None else None, out)
            self.release(part)
            return Act(out, pc.cout)
        if self.head_in_transposed_kernel(pc, hcur) and (res is None or res.c == pc.cout):
            pct = pc
            if res is not None:
                pct = extra.get(key)
                if pct is None:
                    pct = extra[key] = PackedConv([pc.segs[0][0], WSrc(pc.cout, "eye")], list(pc.bias_srcs) or None,
                                                  self.device, pc.dtype)
            (cf,) = self.gn_coeffs([hcur], gamma, beta, groups)
            segs = [(hcur, 9, cf)] + ([(res, 1)] if res is not None else [])
            out = self.conv(segs, pct, head=True)
            self.pool.put(cf)
            return out
        a = self.group_norm([hcur], gamma, beta, True, groups)
        out = self.conv([(a, 9)], pc, res=res, scale=1.0)
        self.release(a)
        return out

    def time_embedding(self, labels, nf, embedding_type, fourier_w, lin0, lin1, P, mods=None):
        """temb MLP (fused kernel) + every block's Dense_0 projection in one launch. Returns (tproj, pitch)."""
        (w0, b0), (w1, b1) = lin0, lin1
        batch = labels.shape[0]
        act_temb = torch.empty(batch, w1.shape[0], device=self.device, dtype=torch.float32)
        self.rec.add(K.time_embedding, labels, nf, embedding_type, fourier_w, w0, b0, w1, b1, act_temb)
        if "dense_w" not in P:
            return None, 0
        tpitch = P["dense_total"]
        tproj = torch.empty(batch, tpitch, device=self.device, dtype=torch.float32)
        self.rec.add(K.dense_rows, act_temb, P["dense_w"], P["dense_b"], tproj)
        return tproj, tpitch

    def dropout(self, a):
        """Dropout_0 of the ResNet blocks: identity outside training (nn.Dropout in eval mode)."""
        return a

    def fir(self, a, mode, taps, add=None, operand=True, norm=None):
        """operand: the result only feeds convolution operands (fp32 plan: rounded to tf32 on store); False for the
        output pyramid, which is added to a head's output. norm: (scale, shift) table of gn_coeffs - the FIR then
        resamples SiLU(GroupNorm(a)) without that tensor ever reaching HBM."""
        b, h, w, p = a.shape
        oh, ow = {"up": (h * 2, w * 2), "down": (h // 2, w // 2), "prefilter": (h + 1, w + 1)}[mode]
        out = self.pool.get((b, oh, ow, p))
        self.rec.add(K.fir_resample, a.t, out, mode, list(taps), add.t if add is not None else None, operand, norm)
        return Act(out, a.c)

    fir_norm = True         # resampling blocks: GroupNorm_0 + SiLU inside the FIR kernel when the sums are known

    def fir_norm_ok(self, srcs):
        a = srcs[0]
        return (self.fir_norm and K.FIR_NORM_DEFAULT and len(srcs) == 1 and a.c % 8 == 0
                and (a.sums is not None or a.pending is not None) and a.t.data_ptr() % 16 == 0)

    def release(self, *acts):
        for a in acts:
            if a is not None:
                self.pool.put(a.t)
                if a.pending is not None:       # nobody asked for its statistics
                    self.pool.put(a.pending[0])
                    a.pending = None

    # -- reference layers ----------------------------------------------------------------------
    def resblock(self, pk, srcs, tproj, tproj_pitch, fir_taps, skip_rescale):
        """ResnetBlockBigGANpp / ResnetBlockDDPMpp forward (models/layerspp.py:195-209,242-274).

        pk: dict with gn0/gn1 (gamma, beta), conv0, conv1 (PackedConv; conv1 already carries the
        skip 1x1 weights as extra K segments when the block has Conv_2 / NIN_0), flags.
        srcs: 1 or 2 Acts (the up path passes [h, skip] instead of materialising torch.cat).
        """
        raw = list(srcs)
        own_raw = False
        resample = pk["up"] or pk["down"]
        temb = tproj[:, pk["temb_off"]:] if tproj is not None else None
        if not resample and self.fusable(srcs, pk["out_ch"]):
            # GroupNorm_0 + SiLU in the prologue of Conv_0: one 9-tap segment per source, no normalised copy
            coefs = self.gn_coeffs(srcs, pk["gn0_w"], pk["gn0_b"], pk["groups0"])
            h1 = self.conv([(a, 9, cf) for a, cf in zip(srcs, coefs)], pk["conv0_split"], temb=temb,
                           temb_pitch=tproj_pitch)
            for cf in coefs:
                self.pool.put(cf)
        else:
            if resample and self.fir_norm_ok(srcs):
                # act(GroupNorm_0(x)) is only ever read by the FIR: normalise the staged tile inside that kernel
                mode = "up" if pk["up"] else "down"
                (cf,) = self.gn_coeffs(srcs, pk["gn0_w"], pk["gn0_b"], pk["groups0"])
                a0 = self.fir(srcs[0], mode, fir_taps, norm=cf)
                self.pool.put(cf)
                raw = [self.fir(srcs[0], mode, fir_taps)]
                own_raw = True
            else:
                a0 = self.group_norm(srcs, pk["gn0_w"], pk["gn0_b"], True, pk["groups0"])
                if resample:
                    mode = "up" if pk["up"] else "down"
                    assert len(srcs) == 1
                    a0r = self.fir(a0, mode, fir_taps)
                    self.release(a0)
                    a0 = a0r
                    raw = [self.fir(srcs[0], mode, fir_taps)]
                    own_raw = True
            h1 = self.conv([(a0, 9)], pk["conv0"], temb=temb, temb_pitch=tproj_pitch)
            self.release(a0)
        scale = SQRT1_2 if skip_rescale else 1.0
        if self.fusable([h1], pk["out_ch"]):
            (cf,) = self.gn_coeffs([h1], pk["gn1_w"], pk["gn1_b"], pk["groups1"])
            first, a1 = (h1, 9, cf), None
        else:
            a1 = self.dropout(self.group_norm([h1], pk["gn1_w"], pk["gn1_b"], True, pk["groups1"]))
            first, cf = (a1, 9), None
        if pk["has_skip_conv"]:
            out = self.conv([first] + [(r, 1) for r in raw], pk["conv1"], scale=scale)
        elif pk["conv1"].identity_skip:
            # x + h as one more K segment with identity weights: exact (bf16 x times 1.0 into the fp32 accumulator)
            assert len(raw) == 1
            out = self.conv([first, (raw[0], 1)], pk["conv1"], scale=scale)
        else:
            assert len(raw) == 1
            out = self.conv([first], pk["conv1"], res=raw[0], scale=scale)
        self.release(h1, a1)
        if cf is not None:
            self.pool.put(cf)
        if own_raw:
            self.release(*raw)
        return out

    def attention(self, pk, x, skip_rescale):
        """AttnBlockpp forward (models/layerspp.py:75-91) as GroupNorm + 5 GEMMs + softmax."""
        b, h, w, _ = x.shape
        c = x.c
        L = h * w
        lp = K.ceil_to(L, 8)
        hn = self.group_norm([x], pk["gn_w"], pk["gn_b"], False, pk["groups"])
        hn_flat = Act(hn.t.view(b, 1, L, hn.pitch), c)
        if self.fused_attention and self.act_dtype == BF16 and x.pitch % 8 == 0 and K.attn_core_supported(L, c):
            # q | k | v in ONE 1x1 GEMM, then QK^T -> softmax -> PV -> projection -> residual in ONE kernel
            # (csrc/attention.cu): 3 launches per block instead of 7, no [B, L, L] tensor in HBM
            qkv = self.pool.get((b, 1, L, 3 * c))
            self.rec.add(K.conv_gemm, [(hn_flat.t, hn.pitch, 0, c, 1)], pk["qkv"].wt, 3 * c, qkv, batch=b, h=1, w=L,
                         n_store=3 * c, n_tile=pk["qkv"].n_tile, bias=pk["qkv"].bias)
            out = self.pool.get((b, h, w, K.ceil_to(c, 8)))
            self.rec.add(K.attn_core, qkv.view(b, L, 3 * c), pk["proj"].wt, pk["proj"].bias,
                         x.t.view(b, L, x.pitch), out.view(b, L, out.shape[-1]), b, L, c,
                         SQRT1_2 if skip_rescale else 1.0)
            self.pool.put(hn.t)
            self.pool.put(qkv)
            return Act(out, c)
        # q | k in one GEMM: [B, L, 2C]
        qk = self.pool.get((b, 1, L, 2 * c))
        self.rec.add(K.conv_gemm, [(hn_flat.t, hn.pitch, 0, c, 1)], pk["qk"].wt, 2 * c, qk, batch=b, h=1, w=L,
                     n_store=2 * c, n_tile=pk["qk"].n_tile, bias=pk["qk"].bias, round_out=True)
        # V^T[b] = Wv^T h[b]^T : A = weight image [C rows, C], B = hn[b] (batched over z)
        vt = self.pool.get((b, c, lp))
        nt_l = K.ceil_to(L, 16) if L <= 256 else K.ceil_to(math.ceil(L / math.ceil(L / 256)), 16)
        self.rec.add(K.conv_gemm, [(pk["wv_img"], c, 0, c, 1)], hn.t, L, vt, batch=1, h=1, w=c, out_pitch=lp,
                     n_store=L, n_tile=nt_l, z_batches=b, a_batch_step=0, wt_batch_stride=L * hn.pitch,
                     wt_pitch=hn.pitch, k_valid=c, wt_rows=L, out_z_stride=c * lp, bias=pk["bv"],
                     bias_per_row=True, round_out=True)
        # logits S[b] = Q[b] K[b]^T (fp32)
        s = self.pool.get((b, L, lp), torch.float32)
        self.rec.add(K.conv_gemm, [(qk, 2 * c, 0, c, 1)], qk, L, s, batch=1, h=1, w=L, out_pitch=lp, n_store=L,
                     n_tile=nt_l, z_batches=b, a_batch_step=1, wt_batch_stride=L * 2 * c, wt_pitch=2 * c,
                     wt_k_off=c, k_valid=c, wt_rows=L, out_z_stride=L * lp)
        p = self.pool.get((b, L, lp))
        self.rec.add(K.softmax_rows, s, p, L, float(int(c) ** (-0.5)))
        # O[b] = P[b] V[b] : A = P (K = L), B = V^T
        o = self.pool.get((b, 1, L, c))
        pc_o_ntile = pk["proj"].n_tile
        self.rec.add(K.conv_gemm, [(p, lp, 0, L, 1)], vt, c, o, batch=1, h=1, w=L, out_pitch=c, n_store=c,
                     n_tile=pc_o_ntile, z_batches=b, a_batch_step=1, wt_batch_stride=c * lp, wt_pitch=lp,
                     k_valid=L, wt_rows=c, out_z_stride=L * c, round_out=True)
        out = self.pool.get((b, h, w, K.ceil_to(c, 8)))
        self.rec.add(K.conv_gemm, [(o, c, 0, c, 1)], pk["proj"].wt, c, out.view(b, 1, L, out.shape[-1]), batch=b,
                     h=1, w=L, n_store=pk["proj"].n_store, n_tile=pk["proj"].n_tile, bias=pk["proj"].bias,
                     res=x.t.view(b, 1, L, x.pitch), res_pitch=x.pitch, scale=SQRT1_2 if skip_rescale else 1.0)
        for t in (hn.t, qk, vt, s, p, o):
            self.pool.put(t)
        return Act(out, c)


# ------------------------------------------------------------------------------------------------
# whole-network plan
# ------------------------------------------------------------------------------------------------
def _gn_params(gn, device):
    return (gn.weight.detach().to(device=device, dtype=torch.float32).contiguous(),
            gn.bias.detach().to(device=device, dtype=torch.float32).contiguous())


class NetEngine:
    """Packs an NCSNpp module's parameters and plans/executes its forward pass on one device."""

    def __init__(self, net):
        self.net = net
        self.precision = "bf16"     # 'bf16' (fast plan) | 'tf32' (fp32 storage + tf32 operands: the reference's class)
        self.device = None
        self.packed = None
        self.param_version = None
        self.param_structure = None
        self._pack_table = None
        self.plans = {}
        self.train_plans = {}

    # -- weights ---------------------------------------------------------------------------------
    def _version(self):
        return tuple((id(p), p._version, p.device) for p in self.net.parameters())

    def _structure(self):
        return tuple((id(p), p.data_ptr(), p.device) for p in self.net.parameters())

    def ensure_packed(self, device, force_refresh=False):
        """Make the packed operands current. Staleness is detected through the parameters' autograd version counters,
        which in-place ops (optimizer steps, `p.copy_`) bump - but writes through `p.data` (the reference's EMA
        `copy_to` / `restore`, models/ema.py:111,149; DDP-style broadcasts) do not. Callers that cannot know how the
        weights were last written (inference entry points: the eval forward, the samplers) pass force_refresh=True: one
        csd_pack_weights launch (0.3 ms for 48.5 M parameters) re-packs from the live parameters unconditionally."""
        v = self._version()
        st = self._structure()
        same_home = (self.packed is not None and self.device == device and self.param_structure == st)
        if same_home and self.param_version == v:
            if force_refresh and not torch.cuda.is_current_stream_capturing():
                self._refresh()
            return
        if device.type != "cuda":
            raise CsdError("the score network runs on CUDA only (libcsd_b200 has no CPU path)")
        if same_home and all(p.device == device for p in self.net.parameters()):
            # same parameter tensors, new values (an optimizer step): re-pack in place, recorded plans stay valid
            self._refresh()
            self.param_version = v
            return
        self.device = device
        self.packed = self._pack(device)
        self.param_version = v
        self.param_structure = st
        self._pack_table = None
        self.plans = {}
        self.train_plans = {}

    @property
    def wt_dtype(self):
        return torch.float32 if self.precision == "tf32" else BF16

    def set_precision(self, precision):
        """Select the arithmetic of every plan built from now on. 'bf16': bf16 activations in HBM and bf16 tensor-core
        operands (fp32 accumulation and statistics) - the fast plan, held to 2e-2 of the output maximum. 'tf32': fp32
        activations in HBM like the reference (sampling/unconditional.py:206, models/ncsnpp.py:264-266) and tf32
        tensor-core operands like its cuDNN convolutions under PyTorch's defaults - held to 1e-3. Drops packed operands
        and plans (they are rebuilt on next use); inference only."""
        if precision not in ("bf16", "tf32"):
            raise ValueError(f"precision {precision!r}: 'bf16' or 'tf32'")
        if precision != self.precision:
            self.precision = precision
            self.packed = None
            self.param_version = None
            self.plans = {}
            self.train_plans = {}

    def invalidate(self):
        """Parameters were modified outside autograd's version tracking (a fused optimizer kernel): re-pack on next use."""
        self.param_version = None

    def _refresh(self):
        """Re-pack every operand from the live parameters: one csd_pack_weights launch when all parameters are fp32
        CUDA tensors (the normal case), the per-tensor torch path otherwise."""
        if self._pack_table is None or self._pack_table[0] != PackedConv.generation:
            self._pack_table = (PackedConv.generation, self._build_pack_table())
        table = self._pack_table[1]
        if table is not None:
            table.run()
            return
        self._refresh_torch()

    def _walk_packed(self, fn):
        seen = set()

        def visit(obj):
            if id(obj) in seen:
                return
            seen.add(id(obj))
            if isinstance(obj, PackedConv):
                fn(obj)
            elif isinstance(obj, dict):
                if "wv_img" in obj:
                    fn(obj)
                for v in list(obj.values()):
                    visit(v)
            elif isinstance(obj, (list, tuple)):
                for v in obj:
                    visit(v)

        visit(self.packed["mods"])
        visit(self.packed.get("extra", {}))

    def _build_pack_table(self):
        jobs, ok = [], [True]

        def add(obj):
            if isinstance(obj, PackedConv):
                j = obj.jobs()
                if j is None:
                    ok[0] = False
                else:
                    jobs.extend(j)
            else:           # attention dict: the V^T GEMM's weight image and row bias
                m = obj["mod"]
                c = m.NIN_2.W.shape[0]
                W, bvec = m.NIN_2.W, m.NIN_2.b
                if not (W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and bvec.is_cuda):
                    ok[0] = False
                    return
                jobs.append(K.pack_job(W.detach(), obj["wv_img"], c, c, 1, c, c, 1, c, 0))
                jobs.append(K.bias_job(obj["bv"], 0, bvec.detach()))

        self._walk_packed(add)
        if "dense_mods" in self.packed:
            off = 0
            for d in self.packed["dense_mods"]:
                if not (d.weight.is_cuda and d.weight.dtype == torch.float32 and d.weight.is_contiguous()):
                    ok[0] = False
                    break
                jobs.append(K.bias_job(self.packed["dense_w"], off * d.weight.shape[1], d.weight.detach()))
                jobs.append(K.bias_job(self.packed["dense_b"], off, d.bias.detach()))
                off += d.weight.shape[0]
        if not ok[0]:
            return None
        return K.PackTable(jobs, self.device)

    def _refresh_torch(self):
        seen = set()

        def visit(obj):
            if id(obj) in seen:
                return
            seen.add(id(obj))
            if isinstance(obj, PackedConv):
                obj.refresh()
            elif isinstance(obj, dict):
                if "wv_img" in obj:
                    self._refresh_attn(obj)
                for v in list(obj.values()):
                    visit(v)
            elif isinstance(obj, (list, tuple)):
                for v in obj:
                    visit(v)

        visit(self.packed["mods"])
        visit(self.packed.get("extra", {}))
        if "dense_mods" in self.packed:
            self._refresh_dense(self.packed)

    @staticmethod
    def _refresh_dense(out):
        total = out["dense_total"] - 512
        dev = out["dense_w"].device
        out["dense_w"][:total] = torch.cat([d.weight.detach().to(dev, torch.float32) for d in out["dense_mods"]], 0)
        out["dense_b"][:total] = torch.cat([d.bias.detach().to(dev, torch.float32) for d in out["dense_mods"]], 0)

    def _pack_resblock(self, m, device, dense_w, dense_b):
        from .models import layers, layerspp
        pk = {"up": getattr(m, "up", False), "down": getattr(m, "down", False)}
        pk["gn0_w"], pk["gn0_b"] = _gn_params(m.GroupNorm_0, device)
        pk["gn1_w"], pk["gn1_b"] = _gn_params(m.GroupNorm_1, device)
        pk["groups0"], pk["groups1"] = m.GroupNorm_0.num_groups, m.GroupNorm_1.num_groups
        pk["mod"] = m
        pk["dtype"] = self.wt_dtype
        pk["conv0"] = PackedConv([WSrc(m.Conv_0.weight)], m.Conv_0.bias, device, self.wt_dtype)
        pk["conv0_w"], pk["conv0_b"] = m.Conv_0.weight, m.Conv_0.bias
        if hasattr(m, "Dense_0"):
            pk["temb_off"] = sum(d.weight.shape[0] for d in dense_w)
            dense_w.append(m.Dense_0)
            dense_b.append(m.Dense_0)
        else:
            pk["temb_off"] = None
        skip_w, skip_b, skip_kind = None, None, "conv"
        if isinstance(m, layerspp.ResnetBlockBigGANpp):
            if hasattr(m, "Conv_2"):
                skip_w, skip_b = m.Conv_2.weight, m.Conv_2.bias
        else:
            if hasattr(m, "NIN_0"):
                skip_w, skip_b, skip_kind = m.NIN_0.W, m.NIN_0.b, "nin"
            elif hasattr(m, "Conv_2"):
                raise CsdError("conv_shortcut=True DDPM ResNet blocks are not supported by the engine")
        pk["has_skip_conv"] = skip_w is not None
        pk["conv1_w"] = m.Conv_1.weight
        pk["conv1_b"] = m.Conv_1.bias
        pk["skip_w"], pk["skip_b"], pk["skip_kind"] = skip_w, skip_b, skip_kind
        pk["skip_cin"] = None if skip_w is None else (skip_w.shape[0] if skip_kind == "nin" else skip_w.shape[1])
        pk["in_ch"], pk["out_ch"] = m.Conv_0.weight.shape[1], m.Conv_0.weight.shape[0]
        return pk

    @staticmethod
    def finish_conv0(pk, split, device):
        """Conv_0 packed with one K segment per raw input source (the fused-prologue form: the concatenation is
        never materialised, so every source is its own 9-tap segment)."""
        if len(split) == 1:
            return pk["conv0"]
        key = ("conv0", tuple(split))
        if key not in pk:
            ws, off = [], 0
            for c in split:
                ws.append(WSrc(pk["conv0_w"], "conv", off, c))
                off += c
            assert off == pk["conv0_w"].shape[1]
            pk[key] = PackedConv(ws, pk["conv0_b"], device, pk["dtype"])
        return pk[key]

    @staticmethod
    def finish_resblock(pk, split, device, identity_skip=False):
        """Build conv1 (+ skip segments). `split`: channel counts of the raw input sources. identity_skip: the
        block has no skip convolution and its convolutions run in the transposed kernel, whose epilogue takes no
        residual tensor - the residual is appended as a 1-tap K segment with identity weights instead."""
        key = ("conv1", tuple(split), identity_skip)
        if key in pk:
            return pk[key]
        if identity_skip and not pk["has_skip_conv"]:
            c = pk["out_ch"]
            assert split == [c] or tuple(split) == (c,)
            pc = PackedConv([WSrc(pk["conv1_w"]), WSrc(c, "eye")], pk["conv1_b"], device, pk["dtype"])
            pc.identity_skip = True
            pk[key] = pc
            return pc
        if pk["has_skip_conv"]:
            ws = [WSrc(pk["conv1_w"])]
            off = 0
            for c in split:
                ws.append(WSrc(pk["skip_w"], pk["skip_kind"], off, c))
                off += c
            assert off == pk["skip_cin"], (off, pk["skip_cin"])
            pc = PackedConv(ws, [(pk["conv1_b"], 0), (pk["skip_b"], 0)], device, pk["dtype"])
        else:
            pc = PackedConv([WSrc(pk["conv1_w"])], pk["conv1_b"], device, pk["dtype"])
        pk[key] = pc
        return pc

    def _pack_attn(self, m, device):
        c = m.NIN_0.W.shape[0]
        pk = {"mod": m}
        pk["gn_w"], pk["gn_b"] = _gn_params(m.GroupNorm_0, device)
        pk["groups"] = m.GroupNorm_0.num_groups
        dt = self.wt_dtype
        pk["qk"] = PackedConv([[WSrc(m.NIN_0.W, "nin"), WSrc(m.NIN_1.W, "nin")]], [(m.NIN_0.b, 0), (m.NIN_1.b, c)], device, dt)
        if dt == BF16:     # fused attention core: q | k | v from one GEMM
            pk["qkv"] = PackedConv([[WSrc(m.NIN_0.W, "nin"), WSrc(m.NIN_1.W, "nin"), WSrc(m.NIN_2.W, "nin")]],
                                   [(m.NIN_0.b, 0), (m.NIN_1.b, c), (m.NIN_2.b, 2 * c)], device, dt)
        # A-operand "image" of the V^T GEMM: rows = output channel, K = input channel
        pk["wv_img"] = torch.empty(1, 1, c, c, device=device, dtype=dt)
        pk["bv"] = torch.zeros(c + 16, device=device, dtype=torch.float32)
        pk["v"] = PackedConv([WSrc(m.NIN_2.W, "nin")], m.NIN_2.b, device, dt)   # training backward (dgrad of V)
        pk["proj"] = PackedConv([WSrc(m.NIN_3.W, "nin")], m.NIN_3.b, device, dt)
        self._refresh_attn(pk)
        return pk

    @staticmethod
    def _refresh_attn(pk):
        m = pk["mod"]
        c = m.NIN_2.W.shape[0]
        w = m.NIN_2.W.detach().t().reshape(1, 1, c, c).to(torch.float32)
        pk["wv_img"].copy_(K.round_tf32(w) if pk["wv_img"].dtype == torch.float32 else w)
        pk["bv"][:c] = m.NIN_2.b.detach()

    def _pack(self, device):
        from .models import layers, layerspp
        net = self.net
        mods = net.all_modules
        packed = [None] * len(mods)
        dense_w, dense_b = [], []
        for i, m in enumerate(mods):
            if isinstance(m, (layerspp.ResnetBlockBigGANpp, layerspp.ResnetBlockDDPMpp, layers.ResnetBlockDDPM)):
                packed[i] = self._pack_resblock(m, device, dense_w, dense_b)
            elif isinstance(m, (layerspp.AttnBlockpp, layers.AttnBlock)):
                packed[i] = self._pack_attn(m, device)
            elif isinstance(m, (layers.Upsample, layers.Downsample)):
                packed[i] = PackedConv([WSrc(m.Conv_0.weight)], m.Conv_0.bias, device, self.wt_dtype) if m.with_conv else None
            elif isinstance(m, torch.nn.Conv2d):
                packed[i] = PackedConv([WSrc(m.weight)], m.bias, device, self.wt_dtype)
            elif isinstance(m, layerspp.Combine):
                packed[i] = PackedConv([WSrc(m.Conv_0.weight)], m.Conv_0.bias, device, self.wt_dtype)
            elif isinstance(m, torch.nn.GroupNorm):
                packed[i] = _gn_params(m, device) + (m.num_groups,)
            elif isinstance(m, (layerspp.Downsample, layerspp.Upsample)):
                if hasattr(m, "Conv2d_0"):
                    packed[i] = PackedConv([WSrc(m.Conv2d_0.weight)], m.Conv2d_0.bias, device, self.wt_dtype)
                elif hasattr(m, "Conv_0"):
                    packed[i] = PackedConv([WSrc(m.Conv_0.weight)], m.Conv_0.bias, device, self.wt_dtype)
            elif isinstance(m, torch.nn.Linear):
                packed[i] = (m.weight.detach().to(device=device, dtype=torch.float32).contiguous(),
                             m.bias.detach().to(device=device, dtype=torch.float32).contiguous())
            elif isinstance(m, layerspp.GaussianFourierProjection):
                packed[i] = m.W.detach().to(device=device, dtype=torch.float32).contiguous()
        out = {"mods": packed}
        if dense_w:
            total = sum(d.weight.shape[0] for d in dense_w)
            pad = 512  # epilogue reads a full n_tile of projection columns past the block's offset
            out["dense_w"] = torch.zeros(total + pad, dense_w[0].weight.shape[1], device=device, dtype=torch.float32)
            out["dense_b"] = torch.zeros(total + pad, device=device, dtype=torch.float32)
            out["dense_total"] = total + pad
            out["dense_mods"] = list(dense_w)
            self._refresh_dense(out)
        return out

    # -- plan ------------------------------------------------------------------------------------
    def plan(self, batch, h, w, c0, c1):
        key = (batch, h, w, c0, c1)
        if key not in self.plans:
            self.plans[key] = NetPlan(self, batch, h, w, c0, c1)
        return self.plans[key]

    def train_plan(self, batch, h, w, c0, c1, want_params=True, want_input=False, dropout=0.0):
        """Forward + backward plan (engine_train.TrainPlan) for differentiating the network."""
        from .engine_train import TrainPlan
        if self.precision != "bf16":
            raise CsdError("the differentiable (training / likelihood) plan exists for the bf16 plan only; "
                           "call set_precision('bf16')")
        key = (batch, h, w, c0, c1, want_params, want_input, float(dropout))
        if key not in self.train_plans:
            self.train_plans[key] = TrainPlan(self, batch, h, w, c0, c1, want_params, want_input, float(dropout))
        return self.train_plans[key]


class NetPlan:
    """Launch list of NCSNpp.forward for one batch shape (models/ncsnpp.py:238-388)."""

    def __init__(self, eng, batch, h, w, c0, c1):
        net, dev = eng.net, eng.device
        self.eng = eng
        self.batch, self.h, self.w = batch, h, w
        self.rec = Recorder()
        self.pool = self._make_pool(dev)
        self.stats = torch.zeros(48 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
        ops = self.ops = self._make_ops(dev)

        # static inputs / outputs
        self.in0 = torch.empty(batch, c0, h, w, device=dev, dtype=torch.float32)
        self.in1 = torch.empty(batch, c1, h, w, device=dev, dtype=torch.float32) if c1 else None
        self.labels = torch.empty(batch, device=dev, dtype=torch.float32)
        self._outs = {}
        self.row_scale = torch.ones(batch, device=dev, dtype=torch.float32)
        self.row_scale1 = torch.ones(batch, device=dev, dtype=torch.float32)

        self.rec.add(self.stats.zero_)
        build = {"ncsnpp": self._build_ncsnpp, "ddpm": self._build_ddpm}[getattr(net, "arch", "ncsnpp")]
        final = build(ops, c0, c1)
        self.final_act = final
        # ---- output: NHWC bf16 -> NCHW fp32, optional per-sample 1/sigma. A network that returns as many channels
        #      as it takes is split back into the (x, y) groups it was fed; otherwise (SR3) there is one output ----
        out_c = net.out_channels
        if c1 and out_c == c0 + c1:
            self.rec.add(K.nhwc_to_nchw, final.t, 0, c0, self._out_view(0, c0), self.row_scale)
            self.rec.add(K.nhwc_to_nchw, final.t, c0, c1, self._out_view(c0, c1), self.row_scale1)
        else:
            self.rec.add(K.nhwc_to_nchw, final.t, 0, out_c, self._out_view(0, out_c), self.row_scale)
        self.graph = None
        self.warm = 0
        self.use_graph = os.environ.get("CSD_NO_GRAPH", "0") != "1"

    def _make_pool(self, dev):
        return BufferPool(dev, self.eng.wt_dtype)

    def _make_ops(self, dev):
        return BlockOps(dev, self.pool, self.rec, self.stats)

    def _time_embedding(self, pk, m_idx, fourier_w):
        """temb MLP + every block's Dense_0 projection in two launches. Returns (tproj, pitch)."""
        net, P = self.eng.net, self.eng.packed
        return self.ops.time_embedding(self.labels, net.nf, net.embedding_type, fourier_w, pk[m_idx], pk[m_idx + 1], P,
                                       (net.all_modules[m_idx], net.all_modules[m_idx + 1]))

    def _input(self, c0, c1):
        """cat(x, y) + `2x - 1` + NCHW fp32 -> NHWC bf16 in one kernel."""
        net, dev = self.eng.net, self.eng.device
        channels = c0 + c1
        xin = Act(torch.empty(self.batch, self.h, self.w, K.ceil_to(channels, 8), device=dev, dtype=self.eng.wt_dtype), channels)
        if net.centered:
            self.rec.add(K.nchw_to_nhwc, self.in0, self.in1, xin.t, 1.0, 0.0)
        else:
            self.rec.add(K.nchw_to_nhwc, self.in0, self.in1, xin.t, 2.0, -1.0)  # x = 2x - 1
        self.xin_act = xin
        return xin

    def _build_ddpm(self, ops, c0, c1):
        """DDPM.forward (models/ddpm.py:150-213)."""
        eng = self.eng
        net, dev, P = eng.net, eng.device, eng.packed
        mods, pk = net.all_modules, P["mods"]
        tproj, tpitch = self._time_embedding(pk, 0, None)
        m_idx = 2
        xin = self._input(c0, c1)
        nearest = (0.0, 1.0, 1.0, 0.0)      # F.interpolate(mode='nearest') x2 as a 2-tap polyphase FIR

        def resblock(idx, srcs):
            p = dict(pk[idx])
            _, sh, sw, _ = srcs[0].shape
            # bf16: the residual rides as an identity K segment (exact); fp32 plan: the tensor core would truncate it to
            # tf32, so it is added in the epilogue instead
            ident = ops.will_transpose(sh, sw, p["out_ch"]) and ops.act_dtype == BF16
            p["conv1"] = eng.finish_resblock(pk[idx], [a.c for a in srcs], dev, identity_skip=ident)
            p["conv0_split"] = eng.finish_conv0(pk[idx], [a.c for a in srcs], dev)
            return ops.resblock(p, srcs, tproj if p["temb_off"] is not None else None, tpitch, None, False)

        refcount = {}

        def hold(a):
            refcount[id(a.t)] = refcount.get(id(a.t), 0) + 1
            return a

        def drop(a):
            refcount[id(a.t)] -= 1
            if refcount[id(a.t)] <= 0:
                ops.release(a)

        hs = [hold(ops.conv([(xin, 9)], pk[m_idx]))]
        m_idx += 1
        for lvl in range(net.num_resolutions):
            for _ in range(net.num_res_blocks):
                hcur = resblock(m_idx, [hs[-1]])
                m_idx += 1
                if hcur.shape[2] in net.attn_resolutions:
                    h2 = ops.attention(pk[m_idx], hcur, False)
                    ops.release(hcur)
                    hcur = h2
                    m_idx += 1
                hs.append(hold(hcur))
            if lvl != net.num_resolutions - 1:
                if not mods[m_idx].with_conv:
                    raise CsdError("DDPM Downsample(with_conv=False) (average pooling) is not supported by the engine")
                a = hs[-1]
                _, ah, aw, _ = a.shape
                # F.pad(x, (0,1,0,1)) + 3x3 stride-2 VALID conv: the bottom/right zeros come from TMA's bounds fill
                hs.append(hold(ops.conv([(a, 9)], pk[m_idx], out_hw=(ah // 2, aw // 2), stride=2, pad=0)))
                m_idx += 1

        hcur = hs[-1]
        hold(hcur)
        h2 = resblock(m_idx, [hcur]); m_idx += 1
        drop(hcur)
        h3 = ops.attention(pk[m_idx], h2, False); m_idx += 1
        ops.release(h2)
        hcur = resblock(m_idx, [h3]); m_idx += 1
        ops.release(h3)

        for lvl in reversed(range(net.num_resolutions)):
            for _ in range(net.num_res_blocks + 1):
                skip = hs.pop()
                h2 = resblock(m_idx, [hcur, skip])
                m_idx += 1
                ops.release(hcur)
                drop(skip)
                hcur = h2
            if hcur.shape[2] in net.attn_resolutions:
                h2 = ops.attention(pk[m_idx], hcur, False)
                ops.release(hcur)
                hcur = h2
                m_idx += 1
            if lvl != 0:
                up = ops.fir(hcur, "up", nearest)
                ops.release(hcur)
                hcur = up
                if mods[m_idx].with_conv:
                    h2 = ops.conv([(hcur, 9)], pk[m_idx])
                    ops.release(hcur)
                    hcur = h2
                m_idx += 1
        assert not hs
        gamma, beta, gn_groups = pk[m_idx]
        m_idx += 1
        pc = pk[m_idx]
        m_idx += 1
        if ops.fusable([hcur], pc.cout):
            (cf,) = ops.gn_coeffs([hcur], gamma, beta, gn_groups)
            final = ops.conv([(hcur, 9, cf)], pc)
        else:
            a = ops.group_norm([hcur], gamma, beta, True, gn_groups)
            final = ops.conv([(a, 9)], pc)
        assert m_idx == len(mods), (m_idx, len(mods))
        return final

    def _build_ncsnpp(self, ops, c0, c1):
        """NCSNpp.forward (models/ncsnpp.py:238-388)."""
        from .models import layerspp
        eng = self.eng
        net, dev, P = eng.net, eng.device, eng.packed
        rec = self.rec
        batch, h, w = self.batch, self.h, self.w
        mods, pk = net.all_modules, P["mods"]
        nf = net.nf
        fir_taps = tuple(net.fir_kernel)
        if len(fir_taps) != 4:
            raise CsdError("the engine supports 4-tap FIR kernels only (config.model.fir_kernel)")
        if not net.fir:
            # fir=False (the DDPM++ configs, e.g. configs/vp/cifar10_ddpmpp_continuous.py): nearest-neighbour x2 up
            # (naive_upsample_2d / F.interpolate, up_or_down_sampling.py:59-63, layerspp.py:116) and 2x2 mean down
            # (naive_downsample_2d / F.avg_pool2d, :66-69, layerspp.py:155) are the same polyphase kernel with taps [1, 1]
            fir_taps = (0.0, 1.0, 1.0, 0.0)
        skip_rescale = net.skip_rescale
        channels = c0 + c1

        m_idx = 0
        # ---- time embedding ----
        fourier_w = None
        if net.embedding_type == "fourier":
            fourier_w = pk[m_idx]
            m_idx += 1
        tproj = None
        tpitch = 0
        if net.conditional:
            tproj, tpitch = self._time_embedding(pk, m_idx, fourier_w)
            m_idx += 2
        # ---- input ----
        cpad = K.ceil_to(channels, 8)
        xin = self.xin_act = Act(torch.empty(batch, h, w, cpad, device=dev, dtype=eng.wt_dtype), channels)
        if net.centered:
            rec.add(K.nchw_to_nhwc, self.in0, self.in1, xin.t, 1.0, 0.0)
        else:
            rec.add(K.nchw_to_nhwc, self.in0, self.in1, xin.t, 2.0, -1.0)  # x = 2x - 1 (ncsnpp.py:264-266)
        input_pyramid = xin if net.progressive_input != "none" else None

        def resblock(idx, srcs):
            p = pk[idx]
            _, sh, sw, _ = srcs[0].shape
            ident = (not (p["up"] or p["down"])) and ops.will_transpose(sh, sw, p["out_ch"]) and ops.act_dtype == BF16
            pc1 = eng.finish_resblock(p, [a.c for a in srcs], dev, identity_skip=ident)
            p = dict(p)
            p["conv1"] = pc1
            p["conv0_split"] = eng.finish_conv0(pk[idx], [a.c for a in srcs], dev)
            return ops.resblock(p, srcs, tproj if p["temb_off"] is not None else None, tpitch, fir_taps, skip_rescale)

        hs = [ops.conv([(xin, 9)], pk[m_idx])]
        m_idx += 1
        refcount = {}

        def hold(a):
            refcount[id(a.t)] = refcount.get(id(a.t), 0) + 1
            return a

        def drop(a):
            n = refcount.get(id(a.t), 0) - 1
            refcount[id(a.t)] = n
            if n <= 0:
                ops.release(a)

        hold(hs[0])
        num_res = net.num_resolutions
        for lvl in range(num_res):
            for _ in range(net.num_res_blocks):
                hcur = resblock(m_idx, [hs[-1]])
                m_idx += 1
                if hcur.shape[2] in net.attn_resolutions:
                    h2 = ops.attention(pk[m_idx], hcur, skip_rescale)
                    ops.release(hcur)
                    hcur = h2
                    m_idx += 1
                hs.append(hold(hcur))
            if lvl != num_res - 1:
                if net.resblock_type == "ddpm":
                    hcur = self._resample_module(ops, mods[m_idx], pk[m_idx], hs[-1], fir_taps, down=True)
                else:
                    hcur = resblock(m_idx, [hs[-1]])
                m_idx += 1
                if net.progressive_input == "input_skip":
                    ip = ops.fir(input_pyramid, "down", fir_taps)
                    if input_pyramid is not xin:
                        ops.release(input_pyramid)
                    input_pyramid = ip
                    if net.combine_method == "sum":
                        # Combine: conv1x1(pyramid) + h (layerspp.py:52-59)
                        h2 = ops.conv([(input_pyramid, 1)], pk[m_idx], res=hcur, scale=1.0)
                    else:
                        # Combine 'cat': cat([conv1x1(pyramid), h], dim=1) - the conv writes the first channels of the
                        # concatenated tensor, h is copied behind them
                        pcc = pk[m_idx]
                        if not ops.fuse_small_gn:      # the training subclass: its tape does not know the copy below
                            raise CsdError("progressive_combine='cat' has no backward plan (inference only)")
                        if pcc.cout % 8 != 0:
                            raise CsdError("progressive_combine='cat' needs a multiple of 8 channels")
                        b_, hh_, ww_, _ = hcur.shape
                        cat = self.pool.get((b_, hh_, ww_, pcc.cout + K.ceil_to(hcur.c, 8)))
                        ops.conv([(input_pyramid, 1)], pcc, out=cat[..., :pcc.cout], out_pitch=cat.shape[-1])
                        rec.add(cat[..., pcc.cout:pcc.cout + hcur.c].copy_, hcur.t[..., :hcur.c])
                        h2 = Act(cat, pcc.cout + hcur.c)
                    ops.release(hcur)
                    hcur = h2
                    m_idx += 1
                elif net.progressive_input == "residual":
                    ip = self._resample_module(ops, mods[m_idx], pk[m_idx], input_pyramid, fir_taps, down=True, res=hcur,
                                               scale=SQRT1_2 if skip_rescale else 1.0)
                    if input_pyramid is not xin:
                        drop(input_pyramid)
                    ops.release(hcur)
                    input_pyramid = hold(ip)   # also pushed on hs below
                    hcur = ip
                    m_idx += 1
                hs.append(hold(hcur))

        hcur = hs[-1]
        hold(hcur)
        h2 = resblock(m_idx, [hcur]); m_idx += 1
        drop(hcur)
        h3 = ops.attention(pk[m_idx], h2, skip_rescale); m_idx += 1
        ops.release(h2)
        hcur = resblock(m_idx, [h3]); m_idx += 1
        ops.release(h3)

        pyramid = None
        for lvl in reversed(range(num_res)):
            for _ in range(net.num_res_blocks + 1):
                skip = hs.pop()
                h2 = resblock(m_idx, [hcur, skip])
                m_idx += 1
                ops.release(hcur)
                drop(skip)
                hcur = h2
            if hcur.shape[2] in net.attn_resolutions:
                h2 = ops.attention(pk[m_idx], hcur, skip_rescale)
                ops.release(hcur)
                hcur = h2
                m_idx += 1
            if net.progressive != "none":
                if not net.fir:
                    raise CsdError("progressive output pyramids with fir=False go through layerspp.Upsample(fir=False), "
                                   "which raises in the reference (models/layerspp.py:116-117); unsupported")
                if net.progressive == "residual":
                    raise CsdError("progressive='residual' relies on upsample_conv_2d, which is dead code in the "
                                   "reference (up_or_down_sampling.py:123 indexes with a negative step); unsupported")
                gn = pk[m_idx]
                m_idx += 1
                extra = self.eng.packed.setdefault("extra", {})
                if lvl == num_res - 1:
                    pyramid = ops.head(gn, pk[m_idx], hcur, extra, m_idx)
                elif ops.head_in_transposed_kernel(pk[m_idx], hcur) and HEAD_MODE == 2:
                    # conv first, then the pyramid's FIR upsampling adds it (fir_nhwc `add`): no residual operand
                    ho = ops.head(gn, pk[m_idx], hcur, extra, m_idx)
                    pu = ops.fir(pyramid, "up", fir_taps, add=ho, operand=False)
                    ops.release(pyramid, ho)
                    pyramid = pu
                else:
                    pu = ops.fir(pyramid, "up", fir_taps, operand=False)
                    ops.release(pyramid)
                    pyramid = ops.head(gn, pk[m_idx], hcur, extra, m_idx, res=pu)
                    ops.release(pu)
                m_idx += 1
            if lvl != 0:
                if net.resblock_type == "ddpm":
                    h2 = self._resample_module(ops, mods[m_idx], pk[m_idx], hcur, fir_taps, down=False)
                else:
                    h2 = resblock(m_idx, [hcur])
                ops.release(hcur)
                hcur = h2
                m_idx += 1
        assert not hs
        if net.progressive == "output_skip":
            final = pyramid
        else:
            gn = pk[m_idx]
            m_idx += 1
            final = ops.head(gn, pk[m_idx], hcur, self.eng.packed.setdefault("extra", {}), m_idx)
            m_idx += 1
        assert m_idx == len(mods), (m_idx, len(mods))
        return final

    def _out_view(self, off, cnt):
        # separate contiguous tensors per output group (the paired model returns a dict of two)
        t = torch.empty(self.batch, cnt, self.h, self.w, device=self.eng.device, dtype=torch.float32)
        self._outs[off] = t
        return t

    def _fir_conv_down(self, ops, pc, a, fir_taps, res=None, scale=1.0):
        """layerspp.Downsample(with_conv=True, fir=True) = conv_downsample_2d + bias
        (up_or_down_sampling.py:144-178): FIR with pad (2,2), then 3x3 stride-2 VALID conv."""
        b, h, w, _ = a.shape
        f = ops.fir(a, "prefilter", fir_taps)
        out = ops.conv([(f, 9)], pc, out_hw=(h // 2, w // 2), res=res, scale=scale, stride=2, pad=0)
        ops.release(f)
        return out

    def _resample_module(self, ops, m, pc, a, fir_taps, down, res=None, scale=1.0):
        """layerspp.Downsample / layerspp.Upsample (models/layerspp.py:94-163) as used by resblock_type='ddpm' and by the
        'residual' input pyramid: fir x with_conv select FIR-filtered or naive resampling, with or without a 3x3 conv.
        `a` is not released (callers own it). res/scale: epilogue residual of the conv forms."""
        b, h, w, _ = a.shape
        if down:
            if m.with_conv and m.fir:          # up_or_down_sampling.Conv2d(down=True) = conv_downsample_2d + bias
                return self._fir_conv_down(ops, pc, a, fir_taps, res=res, scale=scale)
            if m.with_conv:                    # F.pad(x, (0, 1, 0, 1)) + 3x3 stride-2 VALID conv (layerspp.py:151-153)
                return ops.conv([(a, 9)], pc, out_hw=(h // 2, w // 2), res=res, scale=scale, stride=2, pad=0)
            if res is not None:
                raise CsdError("resampling module without convolution cannot take a residual")
            return ops.fir(a, "down", fir_taps, operand=False)      # downsample_2d / avg_pool2d
        if m.with_conv and m.fir:
            raise CsdError("layerspp.Upsample(fir=True, with_conv=True) relies on upsample_conv_2d, which is dead code in "
                           "the reference (up_or_down_sampling.py:123 indexes with a negative step); unsupported")
        if not m.fir:
            raise CsdError("layerspp.Upsample(fir=False) cannot run in the reference: F.interpolate(x, (H*2, W*2), 'nearest') "
                           "(models/layerspp.py:116-117) passes 'nearest' as scale_factor and raises; unsupported")
        up = ops.fir(a, "up", fir_taps, operand=m.with_conv)           # upsample_2d
        if not m.with_conv:
            return up
        out = ops.conv([(up, 9)], pc, res=res, scale=scale)
        ops.release(up)
        return out

    # -- execution ---------------------------------------------------------------------------------
    def run(self):
        self.rec.run()

    def launch(self, reuse_time_embedding=False):
        """Run the launch list: eagerly the first time (and inside someone else's capture), as a
        replayed CUDA graph afterwards. reuse_time_embedding: the labels are the ones of the previous launch (the
        corrector and the predictor of one PC step share vec_t, sampling/unconditional.py:216-220), so the temb MLP and
        the blocks' Dense_0 projections it wrote are still current and their two launches are left out (only where the
        list is run launch by launch - inside a caller's capture; the plan's own graph always holds the whole list)."""
        if reuse_time_embedding and (not self.use_graph or torch.cuda.is_current_stream_capturing()):
            self.rec.run(skip=(K.time_embedding, K.dense_rows))
            return
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            self.rec.run()
            return
        if self.graph is None:
            if self.warm < 1:
                self.rec.run()
                self.warm += 1
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.rec.run()
            self.graph = g
        self.graph.replay()

    def outputs(self):
        return [self._outs[k] for k in sorted(self._outs)]
