"""Training / differentiation plan of the score network: forward launch list that keeps its activations, plus the
reverse launch list that turns output gradients into parameter (and input) gradients.

The reference trains through PyTorch autograd (`loss.backward()` in Lightning's training_step,
lightning_modules/BaseSdeGenerativeModel.py:57-60; losses.py:345-407) and differentiates the score network w.r.t. its
input for the likelihood's Hutchinson divergence (likelihood.py:26-37). Here the same gradients come from a *planned*
backward pass: `TrainOps` records, next to every forward kernel group, what is needed to run its adjoint, and
`TrainPlan` walks that tape in reverse once, emitting libcsd_b200 launches:

  conv / 1x1 / NIN   dgrad  = csd_conv_gemm with flipped-transposed packed weights (zero-stuffed gradient for stride 2),
                     wgrad  = pixel-major re-layout + tcgen05 split-K GEMM + ordered reduce (csrc/wgrad.cu),
                     bias / Dense_0(temb) gradients from per-(image, channel) sums of the output gradient;
  GroupNorm(+SiLU)   three-pass analytic backward over the stored input and its statistics (csrc/backward.cu);
  FIR resampling     the adjoint FIR (up <-> down); attention: softmax backward + 6 batched GEMMs on transposed copies;
  time embedding     small fp32 GEMMs.

Activation gradients are NHWC bf16, parameter gradients fp32 in one flat buffer laid out in `net.parameters()` order
(which is also the DDP all-reduce bucket, distributed.allreduce_gradients). No PyTorch autograd runs inside the network.
"""
import math

import torch

from . import kernels as K
from ._lib import CsdError
from .engine import Act, BlockOps, BufferPool, NetPlan, Recorder, BF16, SQRT1_2


class NoRecyclePool(BufferPool):
    """Activations must survive until the backward pass: nothing is handed back for reuse."""

    def put(self, t):
        pass


class TrainOps(BlockOps):
    """BlockOps that keeps every activation and records a tape of (kind, info) for the backward builder."""

    fuse_small_gn = False   # the GroupNorm backward needs every tensor's channel sums
    ragged_tiles = False    # (the training plan keeps the round-1 kernel choice at the 20 px level)
    fir_norm = False        # the normalised tensor is needed again by the backward pass
    fast_heads = False      # the tape records the plain GroupNorm -> conv (+ epilogue residual) form
    defer_finalize = False  # every activation's channel sums exist as soon as it does (the backward reads them)

    def __init__(self, device, pool, rec, stats_arena, dropout_p=0.0, seed_dev=None):
        super().__init__(device, pool, rec, stats_arena)
        self.tape = []
        self.dropout_p, self.seed_dev, self.n_dropout = dropout_p, seed_dev, 0

    def dropout(self, a):
        if self.dropout_p <= 0.0:
            return a
        self.n_dropout += 1
        salt = self.n_dropout
        self.rec.add(K.dropout, a.t, a.t, self.dropout_p, self.seed_dev, salt)     # in place: only the dropped tensor is used
        self.tape.append(("dropout", dict(act=a, salt=salt)))
        return a

    def fusable(self, srcs, cout):
        # the normalised + activated tensor is the wgrad operand of the convolution, so it has to exist in HBM
        return False

    def release(self, *acts):
        pass

    def conv(self, segs, pc, out_hw=None, temb=None, temb_pitch=0, res=None, scale=1.0, stride=1, pad=1, out=None):
        o = super().conv(segs, pc, out_hw=out_hw, temb=temb, temb_pitch=temb_pitch, res=res, scale=scale, stride=stride,
                         pad=pad, out=out)
        self.tape.append(("conv", dict(segs=[(sg[0], sg[1]) for sg in segs], pc=pc, out=o, temb=temb, res=res,
                                       scale=scale, stride=stride, pad=pad)))
        return o

    def group_norm(self, srcs, gamma, beta, silu, groups=None):
        o = super().group_norm(srcs, gamma, beta, silu, groups)
        c = sum(a.c for a in srcs)
        self.tape.append(("gn", dict(srcs=list(srcs), gamma=gamma, beta=beta, silu=silu,
                                     groups=groups or min(c // 4, 32), out=o)))
        return o

    def fir(self, a, mode, taps, add=None, operand=True):
        o = super().fir(a, mode, taps, add, operand)
        self.tape.append(("fir", dict(src=a, mode=mode, taps=tuple(taps), add=add, out=o)))
        return o

    def time_embedding(self, labels, nf, embedding_type, fourier_w, lin0, lin1, P, mods=None):
        """Decomposed temb MLP (features -> Linear -> SiLU -> Linear -> SiLU) that keeps its intermediates."""
        (w0, b0), (w1, b1) = lin0, lin1
        batch, hid, embed = labels.shape[0], w1.shape[0], w0.shape[1]
        f32 = dict(device=self.device, dtype=torch.float32)
        emb = torch.empty(batch, embed, **f32)
        h0pre, h0 = torch.empty(batch, hid, **f32), torch.empty(batch, hid, **f32)
        tpre, act = torch.empty(batch, hid, **f32), torch.empty(batch, hid, **f32)
        rec = self.rec
        rec.add(K.time_features, labels, nf, embedding_type, fourier_w, emb)
        rec.add(K.sgemm_small, 0, 1, batch, hid, embed, emb, embed, w0, embed, h0pre, hid, bias=b0)
        rec.add(K.silu_f32, h0pre, h0)
        rec.add(K.sgemm_small, 0, 1, batch, hid, hid, h0, hid, w1, hid, tpre, hid, bias=b1)
        rec.add(K.silu_f32, tpre, act)
        info = dict(emb=emb, h0pre=h0pre, h0=h0, tpre=tpre, act=act, w0=w0, b0=b0, w1=w1, b1=b1, tproj=None)
        tproj, tpitch = None, 0
        if "dense_w" in P:
            tpitch = P["dense_total"]
            tproj = torch.empty(batch, tpitch, **f32)
            rec.add(K.dense_rows, act, P["dense_w"], P["dense_b"], tproj)
            info.update(tproj=tproj, tpitch=tpitch, dense_w=P["dense_w"], dense_mods=P["dense_mods"])
        self.tape.append(("temb", info))
        return tproj, tpitch

    def attention(self, pk, x, skip_rescale):
        """AttnBlockpp forward (models/layerspp.py:75-91), same launches as BlockOps.attention, intermediates kept."""
        b, h, w, _ = x.shape
        c = x.c
        L = h * w
        lp = K.ceil_to(L, 8)
        hn = self.group_norm([x], pk["gn_w"], pk["gn_b"], False, pk["groups"])
        qk = self.pool.get((b, 1, L, 2 * c))
        self.rec.add(K.conv_gemm, [(hn.t.view(b, 1, L, hn.pitch), hn.pitch, 0, c, 1)], pk["qk"].wt, 2 * c, qk, batch=b, h=1,
                     w=L, n_store=2 * c, n_tile=pk["qk"].n_tile, bias=pk["qk"].bias)
        vt = self.pool.get((b, c, lp))
        nt_l = K.ceil_to(L, 16) if L <= 256 else K.ceil_to(math.ceil(L / math.ceil(L / 256)), 16)
        self.rec.add(K.conv_gemm, [(pk["wv_img"], c, 0, c, 1)], hn.t, L, vt, batch=1, h=1, w=c, out_pitch=lp,
                     n_store=L, n_tile=nt_l, z_batches=b, a_batch_step=0, wt_batch_stride=L * hn.pitch,
                     wt_pitch=hn.pitch, k_valid=c, wt_rows=L, out_z_stride=c * lp, bias=pk["bv"], bias_per_row=True)
        s = self.pool.get((b, L, lp), torch.float32)
        self.rec.add(K.conv_gemm, [(qk, 2 * c, 0, c, 1)], qk, L, s, batch=1, h=1, w=L, out_pitch=lp, n_store=L,
                     n_tile=nt_l, z_batches=b, a_batch_step=1, wt_batch_stride=L * 2 * c, wt_pitch=2 * c,
                     wt_k_off=c, k_valid=c, wt_rows=L, out_z_stride=L * lp)
        p = self.pool.get((b, L, lp))
        sm_scale = float(int(c) ** (-0.5))
        self.rec.add(K.softmax_rows, s, p, L, sm_scale)
        o = self.pool.get((b, 1, L, c))
        self.rec.add(K.conv_gemm, [(p, lp, 0, L, 1)], vt, c, o, batch=1, h=1, w=L, out_pitch=c, n_store=c,
                     n_tile=pk["proj"].n_tile, z_batches=b, a_batch_step=1, wt_batch_stride=c * lp, wt_pitch=lp,
                     k_valid=L, wt_rows=c, out_z_stride=L * c)
        out = self.pool.get((b, h, w, K.ceil_to(c, 8)))
        scale = SQRT1_2 if skip_rescale else 1.0
        self.rec.add(K.conv_gemm, [(o, c, 0, c, 1)], pk["proj"].wt, c, out.view(b, 1, L, out.shape[-1]), batch=b,
                     h=1, w=L, n_store=pk["proj"].n_store, n_tile=pk["proj"].n_tile, bias=pk["proj"].bias,
                     res=x.t.view(b, 1, L, x.pitch), res_pitch=x.pitch, scale=scale)
        oa = Act(out, c)
        self.tape.append(("attn", dict(pk=pk, x=x, hn=hn, qk=qk, vt=vt, p=p, o=o, out=oa, scale=scale, sm_scale=sm_scale,
                                       L=L, lp=lp, nt_l=nt_l)))
        return oa


class TrainPlan(NetPlan):
    """Forward + backward launch lists of one network for one batch shape."""

    def __init__(self, eng, batch, h, w, c0, c1, want_params=True, want_input=False, dropout=0.0):
        self.want_params, self.want_input = want_params, want_input
        self.dropout_p = dropout
        self.dropout_seed = torch.zeros(1, device=eng.device, dtype=torch.int64)
        super().__init__(eng, batch, h, w, c0, c1)
        self.c0, self.c1 = c0, c1
        self.bwd_graph, self.bwd_warm = None, 0
        self._build_backward()

    def _make_pool(self, dev):
        return NoRecyclePool(dev)

    def _make_ops(self, dev):
        return TrainOps(dev, self.pool, self.rec, self.stats, self.dropout_p, self.dropout_seed)

    # -- gradient storage ------------------------------------------------------------------------------------
    def _param_views(self):
        net, dev = self.eng.net, self.eng.device
        params = list(net.parameters())
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise CsdError("training needs fp32 parameters on the network's CUDA device")
        total = sum(K.ceil_to(p.numel(), 4) for p in params)
        self.gflat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.pviews, off = {}, 0
        self.param_list = params
        self.param_offsets = []
        self._grad_touch, self._cur_entry, self._entry_end = {}, 0, []
        for p in params:
            self.pviews[p.data_ptr()] = self.gflat[off:off + p.numel()].view(p.shape)
            self.param_offsets.append(off)
            off += K.ceil_to(p.numel(), 4)

    def pgrad(self, t):
        """fp32 gradient view of the parameter that owns tensor `t` (a detached alias of the parameter)."""
        v = self.pviews.get(t.data_ptr())
        if v is None:
            raise CsdError("gradient requested for a tensor that is not a network parameter")
        # which tape entry (in backward order) asked for this gradient last: after that entry's launches the gradient
        # is final, which is what lets a data-parallel all-reduce of it start while the rest of the backward still runs
        self._grad_touch[t.data_ptr()] = self._cur_entry
        return v

    def _grad(self, a):
        """(gradient tensor, already initialised?) of an activation (Act or raw tensor)."""
        t = a.t if isinstance(a, Act) else a
        key = t.data_ptr()
        if key not in self.agrads:
            self.agrads[key] = [torch.empty(t.shape, device=t.device, dtype=BF16), False]
        return self.agrads[key]

    def _has_grad(self, a):
        t = a.t if isinstance(a, Act) else a
        e = self.agrads.get(t.data_ptr())
        return e is not None and e[1]

    def _scratch(self, name, numel, dtype):
        cur = self.scratch.get(name)
        if cur is None or cur.numel() < numel:
            self.scratch[name] = torch.empty(numel, device=self.eng.device, dtype=dtype)
        return self.scratch[name]

    def _pixmajor(self, kind, channels, geom, ncopies, sig):
        """Zero-initialised pixel-major buffer, shared by every layer with the same write pattern `sig`."""
        key = (kind, channels, ncopies, geom.splits, geom.row_pitch, geom.wp, geom.q, geom.ips) + tuple(sig)
        if key not in self.pm_bufs:
            self.pm_bufs[key] = K.pixmajor_alloc(geom, channels, ncopies, self.eng.device)
        return self.pm_bufs[key]

    # -- backward builders -------------------------------------------------------------------------------------
    def _build_backward(self):
        dev = self.eng.device
        self.bwd = Recorder()
        self.agrads, self.scratch, self.pm_bufs = {}, {}, {}
        self._param_views()
        self.one = torch.ones(1, 1, device=dev, dtype=torch.float32)
        self.bstats = torch.zeros(32 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
        self.bstats_used = 0
        self.partial_numel = 0
        self._partials = []
        bwd = self.bwd
        bwd.add(self.gflat.zero_)
        bwd.add(self.bstats.zero_)
        tape = self.ops.tape
        self.dtproj = None
        for kind, info in tape:
            if kind == "temb" and info["tproj"] is not None and self.want_params:
                self.dtproj = torch.zeros_like(info["tproj"])
                bwd.add(self.dtproj.zero_)
        # seed: gradient of the NHWC output tensor from the NCHW fp32 output gradients
        final = self.final_act
        self.gouts = [torch.zeros_like(o) for o in self.outputs()]
        gfin = self._grad(final)
        out_c = self.eng.net.out_channels
        if len(self.gouts) == 2:
            bwd.add(K.nchw_grad_to_nhwc, self.gouts[0], self.c0, self.row_scale, self.gouts[1], self.c1, self.row_scale1,
                    gfin[0])
        else:
            bwd.add(K.nchw_grad_to_nhwc, self.gouts[0], out_c, self.row_scale, None, 0, None, gfin[0])
        gfin[1] = True
        self.temb_info = None
        for i, (kind, info) in enumerate(reversed(tape)):
            self._cur_entry = i
            getattr(self, "_bwd_" + kind)(info)
            self._entry_end.append(len(bwd.ops))
        self._cur_entry = len(self._entry_end)          # anything requested from here on is final only at the very end
        # one scratch tensor for the split-K partial sums of every wgrad
        if self.partial_numel:
            part = torch.empty(self.partial_numel, device=dev, dtype=torch.float32)
            for holder in self._partials:
                holder[0] = part
        # input gradient (likelihood divergence): NHWC bf16 -> NCHW fp32, times the d(2x-1)/dx = 2 of the input map
        self.gin = None
        if self.want_input:
            xin = self.xin_act
            if not self._has_grad(xin):
                raise CsdError("no gradient reached the network input")
            in_scale = 1.0 if self.eng.net.centered else 2.0
            self.in_scale_t = torch.full((self.batch,), in_scale, device=dev, dtype=torch.float32)
            self.gin = [torch.empty_like(self.in0)] + ([torch.empty_like(self.in1)] if self.in1 is not None else [])
            g = self._grad(xin)[0]
            bwd.add(K.nhwc_to_nchw, g, 0, self.c0, self.gin[0], self.in_scale_t)
            if self.in1 is not None:
                bwd.add(K.nhwc_to_nchw, g, self.c0, self.c1, self.gin[1], self.in_scale_t)

    def _bstat_slot(self, n):
        n_al = (n + 63) // 64 * 64
        if self.bstats_used + n_al > self.bstats.numel():
            raise CsdError("backward statistics arena too small")
        s = self.bstats[self.bstats_used:self.bstats_used + n]
        self.bstats_used += n_al
        return s

    def _wgrad(self, a_t, a_pitch_c, a_c_off, cin, g_t, g_c_off, cout, taps, geom_hw, stride, pad, scale, dst, strides,
               ci_off, g_pm_cache):
        """dst (+)= scale * wgrad(a, g). a_t: NHWC activation tensor [B, ih, iw, pitch]; g_t: NHWC output gradient.
        geom_hw: extent of the activation grid; g is placed on that grid at (o*stride + 1 - pad) for 3x3 taps."""
        bwd = self.bwd
        b = a_t.shape[0]
        ih, iw = geom_hw
        g_grid = g_pm_cache.get("g_grid", g_t if stride == 1 else None)    # gradient on the activation grid (NHWC)
        if (K.WGRAD_DIRECT and g_grid is not None and ih >= 8 and iw >= 8 and g_grid.shape[-1] >= 64
                and a_t.shape[-1] >= 64 and g_c_off % 8 == 0 and a_c_off % 8 == 0):
            # MN-major operands straight from the NHWC tensors: no pixel-major copies, 3 taps per loaded tile
            splits = K.wgrad_direct_splits(b, ih, iw, cout, cin, taps)
            n = splits * taps * cout * cin
            self.partial_numel = max(self.partial_numel, n)
            holder = [None]
            self._partials.append(holder)

            def run_direct(holder=holder, g_grid=g_grid):
                K.wgrad_direct(g_grid, g_c_off, cout, a_t, a_c_off, cin, taps, holder[0], splits)
                K.wgrad_reduce(holder[0], splits, taps, cout, cin, scale, dst, strides[0], strides[1], strides[2], ci_off,
                               True)

            bwd.add(run_direct)
            return
        geom = K.pixmajor_geometry(b, ih, iw)
        g_off = (1 - pad) if taps == 9 else 0
        gkey = (g_t.data_ptr(), g_c_off, cout, stride, g_off, ih, iw)
        g_pm = g_pm_cache.get(gkey)
        if g_pm is None:
            g_pm = self._pixmajor("g", cout, geom, 1, (g_t.shape[1], g_t.shape[2], stride, g_off))
            bwd.add(K.nhwc_to_pixmajor, g_t, g_c_off, cout, geom, g_pm, stride, g_off)
            g_pm_cache[gkey] = g_pm
        ncop = 3 if taps == 9 else 1
        a_pm = self._pixmajor("a", cin, geom, ncop, (a_t.shape[1], a_t.shape[2], 1, 0))
        bwd.add(K.nhwc_to_pixmajor, a_t, a_c_off, cin, geom, a_pm, 1, 0)
        n = geom.splits * taps * cout * cin
        self.partial_numel = max(self.partial_numel, n)
        holder = [None]
        self._partials.append(holder)

        def run(holder=holder, g_pm=g_pm, a_pm=a_pm, geom=geom):
            K.wgrad_gemm(g_pm, cout, a_pm, cin, taps, geom, holder[0])
            K.wgrad_reduce(holder[0], geom.splits, taps, cout, cin, scale, dst, strides[0], strides[1], strides[2], ci_off,
                           True)

        bwd.add(run)

    def _wgrad_src(self, src, a, g_t, g_c_off, cout, taps, geom_hw, stride, pad, scale, cache):
        """Weight gradient of one WSrc block (rows [g_c_off, g_c_off + cout) of the output) w.r.t. activation `a`."""
        if src.kind == "eye":
            return
        dst = self.pgrad(src.param)
        if src.kind == "nin":           # NIN.W [in, out]
            strides = (1, src.param.shape[1], 0)
        else:                           # nn.Conv2d weight [Cout, Cin, kh, kw]
            kk = src.param.shape[2] * src.param.shape[3]
            strides = (src.param.shape[1] * kk, kk, 1)
        cin = a.c
        self._wgrad(a.t, a.pitch, 0, cin, g_t, g_c_off, cout, taps, geom_hw, stride, pad, scale, dst, strides, src.ci_off,
                    cache)

    def _bwd_conv(self, info):
        out = info["out"]
        if not self._has_grad(out):
            return
        bwd = self.bwd
        pc, scale, stride, pad = info["pc"], info["scale"], info["stride"], info["pad"]
        g = self._grad(out)[0]                      # [B, oh, ow, n_store]
        b, oh, ow, _ = g.shape
        cout = pc.cout
        # ---- bias / Dense_0(temb) gradients from per-(image, channel) sums of g ----
        want_p = self.want_params
        temb = info["temb"]
        if self.dtproj is None:
            temb = None
        if want_p and (pc.bias_srcs or temb is not None):
            c8 = g.shape[-1]                      # channel pitch: padding columns of g are zero
            sums = self._bstat_slot(b * c8 * 2).view(b, c8, 2)
            bwd.add(K.gn_chan_stats, g, c8, sums)
            db = [self.pgrad(bp) for bp, off in pc.bias_srcs] if want_p else []
            if any(off != 0 for _, off in pc.bias_srcs):
                raise CsdError("stacked bias sources are handled by the attention backward only")
            dt = None
            if temb is not None:
                dt = self.dtproj[:, temb.storage_offset():]
            bwd.add(K.bias_temb_grad, sums, cout, scale, db[0] if len(db) > 0 else None, db[1] if len(db) > 1 else None,
                    dt, self.dtproj.shape[1] if dt is not None else 0)
        # ---- residual ----
        res = info["res"]
        if res is not None:
            gr = self._grad(res)
            bwd.add(K.axpy, g, gr[0], scale, gr[1])
            gr[1] = True
        # ---- per segment: data gradient and weight gradient ----
        a0 = info["segs"][0][0]
        ih, iw = a0.shape[1], a0.shape[2]
        g_in = g                                    # gradient on the input grid (zero-stuffed for stride 2)
        if stride != 1:
            g_in = self.pool.get((b, ih, iw, g.shape[-1]))
            bwd.add(K.zero_stuff, g, g_in, stride, 1 - pad)
        cache = {"g_grid": g_in}
        for i, (a, taps) in enumerate(info["segs"]):
            srcs = pc.segs[i]
            is_input = a is self.xin_act
            need_dgrad = (not is_input) or self.want_input
            if need_dgrad:
                ga = self._grad(a)
                if len(srcs) == 1 and srcs[0].kind == "eye":
                    bwd.add(K.axpy, g, ga[0], scale, ga[1])
                else:
                    dp = pc.dgrad(i, scale)
                    gsrc = g_in if taps == 9 else g
                    if taps != 9 and stride != 1:
                        raise CsdError("strided 1x1 convolutions are not part of the networks")
                    bwd.add(K.conv_gemm, [(gsrc, gsrc.shape[-1], 0, cout, taps)], dp.wt, dp.cout, ga[0], batch=b,
                            h=a.shape[1], w=a.shape[2], n_store=ga[0].shape[-1], n_tile=dp.n_tile,
                            res=ga[0] if ga[1] else None, res_pitch=ga[0].shape[-1] if ga[1] else 0,
                            transposed=None if not ga[1] else False)
                ga[1] = True
            if want_p:
                co = 0
                for src in srcs:
                    rows = src.weight(self.eng.device).shape[0]
                    self._wgrad_src(src, a, g, co, rows, taps, (ih, iw), stride, pad, scale, cache)
                    co += rows

    def _bwd_gn(self, info):
        out = info["out"]
        if not self._has_grad(out):
            return
        bwd = self.bwd
        dy = self._grad(out)[0]
        srcs = info["srcs"]
        s0 = srcs[0]
        s1 = srcs[1] if len(srcs) > 1 else None
        b, h, w, _ = s0.shape
        hw = h * w
        C = sum(a.c for a in srcs)
        f32 = dict(device=self.eng.device, dtype=torch.float32)
        coef0 = torch.empty(b, s0.c, 2, **f32)
        coef1 = torch.empty(b, s1.c, 2, **f32) if s1 is not None else None
        bwd.add(K.gn_coeffs, s0.sums, s0.c, s1.sums if s1 else None, s1.c if s1 else 0, info["gamma"], info["beta"], coef0,
                coef1, hw, info["groups"], 1e-6)
        s = self._bstat_slot(b * C * 2).view(b, C, 2)
        bwd.add(K.gn_bwd_stats, s0.t, s0.c, dy, 0, coef0, s, 0, info["silu"])
        if s1 is not None:
            bwd.add(K.gn_bwd_stats, s1.t, s1.c, dy, s0.c, coef1, s, s0.c, info["silu"])
        bcoef = torch.empty(b, C, 4, **f32)
        dg = self.pgrad(info["gamma"]) if self.want_params else None
        dbt = self.pgrad(info["beta"]) if self.want_params else None
        bwd.add(K.gn_bwd_coeffs, s0.sums, s0.c, s1.sums if s1 else None, s1.c if s1 else 0, info["gamma"], s, bcoef, dg, dbt,
                hw, info["groups"], 1e-6)
        off = 0
        for a, cf in ((s0, coef0), (s1, coef1)):
            if a is None:
                continue
            if a is self.xin_act and not self.want_input:
                off += a.c
                continue
            ga = self._grad(a)
            bwd.add(K.gn_bwd_apply, a.t, a.c, dy, off, cf, bcoef, off, ga[0], info["silu"], ga[1])
            ga[1] = True
            off += a.c

    def _bwd_dropout(self, info):
        a = info["act"]
        if not self._has_grad(a):
            return
        g = self._grad(a)[0]
        self.bwd.add(K.dropout, g, g, self.dropout_p, self.dropout_seed, info["salt"])

    def _bwd_fir(self, info):
        out = info["out"]
        if not self._has_grad(out):
            return
        bwd = self.bwd
        g = self._grad(out)[0]
        src = info["src"]
        if not (src is self.xin_act and not self.want_input):
            gs = self._grad(src)
            bwd.add(K.fir_resample_bwd, g, gs[0], info["mode"], info["taps"], gs[1])
            gs[1] = True
        add = info["add"]
        if add is not None:
            ga = self._grad(add)
            bwd.add(K.axpy, g, ga[0], 1.0, ga[1])
            ga[1] = True

    def _bwd_temb(self, info):
        """Runs last (the time embedding is the first forward op): dtproj has been accumulated by every block."""
        if info["tproj"] is None or not self.want_params:
            return
        bwd = self.bwd
        dt = self.dtproj
        b, tpitch = dt.shape
        hid = info["act"].shape[1]
        embed = info["emb"].shape[1]
        f32 = dict(device=self.eng.device, dtype=torch.float32)
        ones = torch.ones(1, b, **f32)
        total = sum(d.weight.shape[0] for d in info["dense_mods"])
        # every block's Dense_0 gradient in two GEMMs over the concatenated projection (dWcat = dt^T act, dbcat = colsum dt),
        # then one batched scatter (csd_pack_weights fp32 jobs) into the per-block parameter gradients
        dwcat = torch.empty(total, hid, **f32)
        dbcat = torch.empty(total, **f32)
        bwd.add(K.sgemm_small, 1, 0, total, hid, b, dt, tpitch, info["act"], hid, dwcat, hid)
        bwd.add(K.sgemm_small, 0, 0, 1, total, b, ones, b, dt, tpitch, dbcat, total)
        jobs, off = [], 0
        for d in info["dense_mods"]:
            n = d.weight.shape[0]
            gw, gb = self.pgrad(d.weight), self.pgrad(d.bias)
            jobs.append(K.bias_job(gw, 0, dwcat[off:off + n], gw))       # gw += dwcat rows (kind 1: dst = src + src2)
            jobs.append(K.bias_job(gb, 0, dbcat[off:off + n], gb))
            off += n
        self.dense_scatter = K.PackTable(jobs, self.eng.device)
        bwd.add(self.dense_scatter.run)
        dact = torch.empty(b, hid, **f32)
        bwd.add(K.sgemm_small, 0, 0, b, hid, total, dt, tpitch, info["dense_w"], hid, dact, hid)
        dtpre = torch.empty(b, hid, **f32)
        bwd.add(K.silu_f32, info["tpre"], dtpre, dact)
        bwd.add(K.sgemm_small, 1, 0, hid, hid, b, dtpre, hid, info["h0"], hid, self.pgrad(info["w1"]), hid, beta=1.0)
        bwd.add(K.sgemm_small, 0, 0, 1, hid, b, ones, b, dtpre, hid, self.pgrad(info["b1"]), hid, beta=1.0)
        dh0 = torch.empty(b, hid, **f32)
        bwd.add(K.sgemm_small, 0, 0, b, hid, hid, dtpre, hid, info["w1"], hid, dh0, hid)
        dh0pre = torch.empty(b, hid, **f32)
        bwd.add(K.silu_f32, info["h0pre"], dh0pre, dh0)
        bwd.add(K.sgemm_small, 1, 0, hid, embed, b, dh0pre, hid, info["emb"], embed, self.pgrad(info["w0"]), embed, beta=1.0)
        bwd.add(K.sgemm_small, 0, 0, 1, hid, b, ones, b, dh0pre, hid, self.pgrad(info["b0"]), hid, beta=1.0)

    def _bwd_attn(self, info):
        out = info["out"]
        if not self._has_grad(out):
            return
        bwd, pool, dev = self.bwd, self.pool, self.eng.device
        pk, x, hn = info["pk"], info["x"], info["hn"]
        b, h, w, _ = x.shape
        c, L, lp, nt_l = x.c, info["L"], info["lp"], info["nt_l"]
        scale = info["scale"]
        m = pk["mod"]
        g = self._grad(out)[0].view(b, 1, L, -1)                 # gradient of the block output
        gx = self._grad(x)
        cache = {}
        # residual: out = scale * (proj(o) + x)
        bwd.add(K.axpy, g, gx[0], scale, gx[1])
        gx[1] = True
        # ---- output projection NIN_3 ----
        o = info["o"]                                           # [B, 1, L, c]
        if self.want_params:
            sums = self._bstat_slot(b * c * 2).view(b, c, 2)
            bwd.add(K.gn_chan_stats, g, c, sums)
            bwd.add(K.bias_temb_grad, sums, c, scale, self.pgrad(m.NIN_3.b), None, None, 0)
            self._wgrad(o, c, 0, c, g, 0, c, 1, (1, L), 1, 0, scale, self.pgrad(m.NIN_3.W), (1, c, 0), 0, cache)
        dpj = pk["proj"].dgrad(0, scale)
        do = pool.get((b, 1, L, c))
        bwd.add(K.conv_gemm, [(g, g.shape[-1], 0, c, 1)], dpj.wt, c, do, batch=b, h=1, w=L, n_store=c, n_tile=dpj.n_tile)
        # ---- O = P V ----
        vt, p, qk = info["vt"], info["p"], info["qk"]
        v = pool.get((b, L, c))
        bwd.add(K.transpose, vt, v, c, L)                        # V [B, L, c]
        dp = pool.get((b, L, lp), torch.float32)
        bwd.add(K.conv_gemm, [(do, c, 0, c, 1)], v, L, dp, batch=1, h=1, w=L, out_pitch=lp, n_store=L, n_tile=nt_l,
                z_batches=b, a_batch_step=1, wt_batch_stride=L * c, wt_pitch=c, k_valid=c, wt_rows=L, out_z_stride=L * lp)
        pt = pool.get((b, L, lp))
        bwd.add(K.transpose, p, pt, L, L)                        # P^T [B, L', lp]
        dot = pool.get((b, c, lp))
        bwd.add(K.transpose, do.view(b, L, c), dot, L, c)        # dO^T [B, c, lp]
        dv = pool.get((b, 1, L, c))
        bwd.add(K.conv_gemm, [(pt, lp, 0, L, 1)], dot, c, dv, batch=1, h=1, w=L, out_pitch=c, n_store=c,
                n_tile=pk["proj"].n_tile, z_batches=b, a_batch_step=1, wt_batch_stride=c * lp, wt_pitch=lp, k_valid=L,
                wt_rows=c, out_z_stride=L * c)
        # ---- softmax ----
        ds = pool.get((b, L, lp))
        bwd.add(K.softmax_bwd, p, dp, ds, L, info["sm_scale"])
        # ---- S = Q K^T ----
        qk3 = qk.view(b, L, 2 * c)
        kt = pool.get((b, c, lp))
        qt = pool.get((b, c, lp))
        bwd.add(self._transpose_cols, qk3, c, c, kt, L)          # K^T [B, c, lp]
        bwd.add(self._transpose_cols, qk3, 0, c, qt, L)          # Q^T
        dst = pool.get((b, L, lp))
        bwd.add(K.transpose, ds, dst, L, L)                      # dS^T
        dqk = pool.get((b, 1, L, 2 * c))
        # dQ = dS K  -> columns [0, c) of dqk ; dK = dS^T Q -> columns [c, 2c)
        bwd.add(K.conv_gemm, [(ds, lp, 0, L, 1)], kt, c, dqk, batch=1, h=1, w=L, out_pitch=2 * c, n_store=c,
                n_tile=pk["proj"].n_tile, z_batches=b, a_batch_step=1, wt_batch_stride=c * lp, wt_pitch=lp, k_valid=L,
                wt_rows=c, out_z_stride=L * 2 * c)
        bwd.add(K.conv_gemm, [(dst, lp, 0, L, 1)], qt, c, dqk.view(-1)[c:], batch=1, h=1, w=L, out_pitch=2 * c, n_store=c,
                n_tile=pk["proj"].n_tile, z_batches=b, a_batch_step=1, wt_batch_stride=c * lp, wt_pitch=lp, k_valid=L,
                wt_rows=c, out_z_stride=L * 2 * c)
        # ---- q | k | v projections (NIN_0..2) ----
        hn4 = hn.t.view(b, 1, L, hn.pitch)
        if self.want_params:
            sums = self._bstat_slot(b * 2 * c * 2).view(b, 2 * c, 2)
            bwd.add(K.gn_chan_stats, dqk, 2 * c, sums)
            dbqk = self._bstat_slot(2 * c)
            bwd.add(K.bias_temb_grad, sums, 2 * c, 1.0, dbqk, None, None, 0)
            bwd.add(self._add_f32, self.one, dbqk, 0, c, self.pgrad(m.NIN_0.b))
            bwd.add(self._add_f32, self.one, dbqk, c, c, self.pgrad(m.NIN_1.b))
            sums_v = self._bstat_slot(b * c * 2).view(b, c, 2)
            bwd.add(K.gn_chan_stats, dv, c, sums_v)
            bwd.add(K.bias_temb_grad, sums_v, c, 1.0, self.pgrad(m.NIN_2.b), None, None, 0)
            hn_act = Act(hn4, c)
            self._wgrad(hn4, hn.pitch, 0, c, dqk, 0, c, 1, (1, L), 1, 0, 1.0, self.pgrad(m.NIN_0.W), (1, c, 0), 0, cache)
            self._wgrad(hn4, hn.pitch, 0, c, dqk, c, c, 1, (1, L), 1, 0, 1.0, self.pgrad(m.NIN_1.W), (1, c, 0), 0, cache)
            self._wgrad(hn4, hn.pitch, 0, c, dv, 0, c, 1, (1, L), 1, 0, 1.0, self.pgrad(m.NIN_2.W), (1, c, 0), 0, cache)
        ghn = self._grad(hn)
        dqkp = pk["qk"].dgrad(0, 1.0)
        ghn4 = ghn[0].view(b, 1, L, -1)
        bwd.add(K.conv_gemm, [(dqk, 2 * c, 0, 2 * c, 1)], dqkp.wt, c, ghn4, batch=b, h=1, w=L, n_store=ghn4.shape[-1],
                n_tile=dqkp.n_tile)
        dvp = pk["v"].dgrad(0, 1.0)
        bwd.add(K.conv_gemm, [(dv, c, 0, c, 1)], dvp.wt, c, ghn4, batch=b, h=1, w=L, n_store=ghn4.shape[-1],
                n_tile=dvp.n_tile, res=ghn4, res_pitch=ghn4.shape[-1])
        ghn[1] = True
        # the GroupNorm that produced hn is an ordinary tape entry recorded before this one: it runs next

    @staticmethod
    def _transpose_cols(src3, col_off, cols, out, rows):
        """out[z, c, r] = src3[z, r, col_off + c] for a [z, rows, pitch] tensor."""
        from ._lib import lib, check
        import ctypes
        z, _, pitch = src3.shape
        esz = 2
        check(lib().csd_transpose_bf16(ctypes.c_void_p(src3.data_ptr() + col_off * esz), pitch, src3.shape[1] * pitch,
                                       ctypes.c_void_p(out.data_ptr()), out.shape[-1], out.shape[-2] * out.shape[-1], rows,
                                       cols, z, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    @staticmethod
    def _add_f32(one, src, off, n, dst):
        """dst[0:n] += src[off:off+n] through the small-GEMM entry point (1x1 'GEMM' with beta = 1)."""
        K.sgemm_small(0, 0, 1, n, 1, one, 1, src[off:], n, dst, n, beta=1.0)

    # -- execution -----------------------------------------------------------------------------------------------
    # -- data-parallel overlap: the reverse list in segments, finished gradient ranges handed to a sync hook ----------
    def grad_segments(self, nseg):
        """Split the reverse launch list at tape-entry boundaries into <= nseg segments of roughly equal launch counts.
        Returns [(op_lo, op_hi, ranges)]: `ranges` = coalesced (offset, length) spans of the flat gradient buffer whose
        last writer sits inside the segment - they are final once the segment has run. Every element of the buffer
        appears in exactly one segment's ranges."""
        key = ("segs", nseg)
        if key in self.scratch:
            return self.scratch[key]
        n_ops = len(self.bwd.ops)
        ends = self._entry_end
        cuts = []
        for k in range(1, nseg):
            target = n_ops * k / nseg
            e = min(range(len(ends)), key=lambda i: abs(ends[i] - target))
            if not cuts or e > cuts[-1]:
                cuts.append(e)
        bounds = [(-1, 0)] + [(e, ends[e]) for e in cuts] + [(len(ends) + 1, n_ops)]
        touch = self._grad_touch
        segs = []
        for (e_lo, op_lo), (e_hi, op_hi) in zip(bounds[:-1], bounds[1:]):
            spans = []
            for p, off in zip(self.param_list, self.param_offsets):
                last = touch.get(p.data_ptr(), len(ends) + 1)       # never requested: stays zero, "final" at the end
                if e_lo < last <= e_hi:
                    spans.append((off, K.ceil_to(p.numel(), 4)))
            spans.sort()
            merged = []
            for off, n in spans:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + n)
                else:
                    merged.append((off, n))
            segs.append((op_lo, op_hi, merged))
        self.scratch[key] = segs
        return segs

    def _run_segmented(self, sync):
        segs = self.grad_segments(sync.segments)
        capturing = torch.cuda.is_current_stream_capturing()
        if self.use_graph and not capturing and self.bwd_warm >= 1 and not hasattr(self, "_seg_graphs"):
            torch.cuda.synchronize()
            self._seg_graphs = []
            for lo, hi, _ in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for fn, a, kw in self.bwd.ops[lo:hi]:
                        fn(*a, **kw)
                self._seg_graphs.append(g)
        graphs = getattr(self, "_seg_graphs", None) if (self.use_graph and not capturing) else None
        for i, (lo, hi, ranges) in enumerate(segs):
            if graphs is not None:
                graphs[i].replay()
            else:
                for fn, a, kw in self.bwd.ops[lo:hi]:
                    fn(*a, **kw)
            sync(self.gflat, ranges, last=(i == len(segs) - 1))
        self.bwd_warm += 1

    def run_backward(self):
        """Run the reverse launch list: eagerly the first time, as a replayed CUDA graph afterwards (the list is static:
        every buffer, including the parameter-gradient buffer and the split-K scratch, belongs to the plan). With a
        gradient-sync hook on the engine (distributed.enable_gradient_overlap) the list runs in segments and every
        segment's finished gradient ranges are all-reduced on a side stream while the next segment computes."""
        sync = getattr(self.eng, "grad_sync", None)
        if sync is not None and self.want_params:
            self._run_segmented(sync)
            return
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            self.bwd.run()
            return
        if self.bwd_graph is None:
            if self.bwd_warm < 1:
                self.bwd.run()
                self.bwd_warm += 1
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.bwd.run()
            self.bwd_graph = g
        self.bwd_graph.replay()

    def param_grads(self):
        """Fresh fp32 gradient tensors in net.parameters() order (views of one clone of the flat buffer)."""
        flat = self.gflat.clone()
        return [flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.param_offsets, self.param_list)]
