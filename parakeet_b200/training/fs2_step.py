"""FastSpeech2 training step on B200 (reference: FastSpeech2Updater.update_core, parakeet/models/fastspeech2/
fastspeech2_updater.py:51-99; data-parallel set-up examples/fastspeech2/train.py:51-56,135-139).

    forward (train mode: BatchNorm uses batch statistics; Dropout at the reference's sites with Philox masks that the backward
    pass regenerates from (seed, step, site) - pk_dropout - so no mask is ever stored)
    -> FastSpeech2Loss (use_masking=True) -> backward -> mean all-reduce of the gradients over ranks (DataParallel)
    -> paddle.optimizer.Adam step.

All parameters live in ONE flat fp32 buffer (the model's state-dict entries are views into it), all gradients in a second
flat buffer: the data-parallel exchange is a single NCCL all-reduce of that buffer per step over NVLink, the 1/world
scale is folded into the fused Adam kernel (pk_adam).  Every FLOP runs in libparakeet_b200.so: GEMM-shaped gradients
(dgrad = conv with flipped taps, wgrad = dY^T X over the flattened batch x time axis, attention dQ/dK/dV/dP) reuse
pk_conv_gemm on transposed split planes (pk_transpose_planes); the rest are the row-wise kernels of train.cu.
torch is used for buffers, views, permutes / copies (layout plumbing) and torch.distributed.
"""
import ctypes as C
import math
import os

import torch
import torch.distributed as dist

from .. import _lib, ops
from ..models.fastspeech2 import FastSpeech2, _i32
from ..ops import Split, _ptr, _stream
from ..graph import GraphRunner
from . import wgrad
from .flat import FlatBuffers

BUFFERS = ("_mean", "_variance")


def _ceil64(n):
    return (n + 63) // 64 * 64


def pack_dev(w):
    """[n, k, taps] (or [n, k]) fp32 CUDA tensor -> K-major split planes [n, taps * Kp] (device-side ops.pack_weight)."""
    if w.dim() == 2:
        w = w.unsqueeze(-1)
    n, k, taps = w.shape
    kp = _ceil64(k)
    packed = torch.zeros(n, taps, kp, dtype=torch.float32, device=w.device)
    packed[:, :, :k] = w.permute(0, 2, 1)
    return Split.from_f32(packed.reshape(n, taps * kp))


class FastSpeech2TrainStep:
    def __init__(self, model: FastSpeech2, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8,
                 stop_gradient_from_pitch_predictor=None, stop_gradient_from_energy_predictor=None, process_group=None,
                 dropout=True, seed=0, use_graphs=None):
        """dropout: True -> the model's constructor rates (the reference trains in model.train() mode), a dict of the
        reference's rate keywords to override them, or False / None -> every rate 0 (deterministic step, parity tests).
        seed: base seed of the Philox masks; every rank should pass its own (paddle seeds each process's generator).
        use_graphs: replay forward + backward as ONE CUDA graph per batch shape (eager the first time a shape is seen, captured
        the second, replayed afterwards); None -> env PK_TRAIN_GRAPH (default on).  The step is host-bound otherwise (~700
        launches of 10-40 us); bucketing samplers repeat shapes, so do synthetic benchmarks."""
        if not model.device.type == "cuda":
            raise _lib.PkError("training needs a CUDA device (no CPU fallback)")
        if model.spk_embed_dim is not None or model.tone_embed_dim is not None:
            raise NotImplementedError("the training step covers the single-speaker recipe (speaker / tone conditioning is inference-only)")
        self.m = model
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.sg_pitch = model.stop_gradient_from_pitch_predictor if stop_gradient_from_pitch_predictor is None else stop_gradient_from_pitch_predictor
        self.sg_energy = model.stop_gradient_from_energy_predictor if stop_gradient_from_energy_predictor is None else stop_gradient_from_energy_predictor
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        dev = model.device
        self.dev = dev
        self.overlap = os.environ.get("PK_TRAIN_OVERLAP", "1") != "0"      # parameter gradients on a side stream (on_side)
        self._side, self._side_used, self._keep = None, False, []
        names = [k for k in model._params if not k.endswith(BUFFERS)]
        self.buffers = FlatBuffers(model._params, names, dev)      # the model's tensors become views of one flat buffer
        self.flat, self.gflat, self.grads = self.buffers.flat, self.buffers.gflat, self.buffers.grads
        self.adam_m = torch.zeros(self.buffers.total, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(self.buffers.total, dtype=torch.float32, device=dev)
        model._packed = None
        self.step_count = 0
        if dropout is True:
            self.rates = dict(model.dropout_rates)
        elif isinstance(dropout, dict):
            self.rates = {**model.dropout_rates, **dropout}
        else:
            self.rates = {k: 0.0 for k in model.dropout_rates}
        self.seed = int(seed)
        # completed steps, on the device: the dropout kernels add it to their step argument, so a captured graph of forward +
        # backward draws new masks on every replay
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._fb_graphs = GraphRunner(max_graphs=16)
        self._zp = wgrad.ZeroPlanes(max_geoms=16, on_evict=self._fb_graphs.drop)     # planes and graph of a batch shape go together
        self.use_graphs = (os.environ.get("PK_TRAIN_GRAPH", "1") != "0") if use_graphs is None else bool(use_graphs)
        # workspace of the BatchNorm / LayerNorm reductions: 2 floats per channel (pk_batch_norm_train / _bwd)
        widest = max([model.odim, model.adim] + [int(v.shape[0]) for k, v in model._params.items() if k.startswith("postnet.")])
        self.sums = torch.zeros(max(4096, 2 * widest), dtype=torch.float32, device=dev)
        if model.adim > 512:
            raise NotImplementedError("pk_layer_norm_bwd supports rows of at most 512 channels (adim)")
        if self.world > 1:
            # paddle.DataParallel broadcasts rank 0's parameters and buffers at construction (train.py:117-119)
            dist.broadcast(self.flat, src=0, group=process_group)
            for k, v in model._params.items():
                if k.endswith(BUFFERS):
                    dist.broadcast(v, src=0, group=process_group)

    # ------------------------------------------------------------------------------------------------------------
    # GEMM-shaped forward / backward pieces
    # ------------------------------------------------------------------------------------------------------------
    def P(self, name):
        return self.m._params[name]

    def _pack(self, key, fn):
        v = self._packs.get(key)
        if v is None:
            v = self._packs[key] = fn()
        return v

    @staticmethod
    def site(stack, layer, kind):
        """stack: 0 encoder, 1 decoder, 2 pitch, 3 energy, 4 duration predictor, 5 postnet; kind: 0 positional encoding,
        1 attention probabilities, 2 attention sub-layer output, 3 feed-forward hidden, 4 feed-forward sub-layer output,
        5 predictor layer, 6 postnet layer; (6, 0, 7) / (6, 0, 8): pitch / energy embedding
        (oracle/fastspeech2.py: dropout_site restates this numbering)."""
        return stack * 1000 + layer * 10 + kind

    def drop(self, x, p, site, **kw):
        """Forward AND backward: the mask depends only on (seed, step, site, element index)."""
        return ops.dropout(x, p, self.seed, site, 1, step_dev=self.step_dev, **kw)      # step = 1 + completed steps (device counter)

    def w_fwd(self, name, kind):
        w = self.P(name)
        return self._pack(("f", name), lambda: pack_dev(w.t().contiguous() if kind == "lin" else w))

    def w_bwd(self, name, kind):
        w = self.P(name)
        return self._pack(("b", name), lambda: pack_dev(w if kind == "lin" else w.flip(-1).permute(1, 0, 2).contiguous()))

    def dims(self, name, kind):
        w = self.P(name)
        return (w.shape[0], w.shape[1], 1) if kind == "lin" else (w.shape[1], w.shape[0], w.shape[2])   # (cin, cout, taps)

    def layer_fwd(self, x, wname, bname, kind, act=None, residual=None, f32=True, split=False):
        cin, cout, taps = self.dims(wname, kind)
        return ops.conv_gemm(x, self.w_fwd(wname, kind), n=cout, k=cin, taps=taps, bias=self.P(bname) if bname else None, act=act,
                             residual=residual, out_f32=f32, out_split=split)

    def layer_bwd(self, dy, x_saved, wname, bname, kind, need_dx=True):
        """dy fp32 (B,T,cout), x_saved split (B,T,cin): accumulates grads of weight / bias, returns dx fp32 (B,T,cin)."""
        cin, cout, taps = self.dims(wname, kind)
        B, T = dy.shape[0], dy.shape[1]
        if cout % 8:
            dy8 = torch.zeros(B, T, (cout + 7) // 8 * 8, dtype=torch.float32, device=dy.device)   # TMA row pitch: 16 bytes
            dy8[..., :cout] = dy
            dys = Split.from_f32(dy8)
        else:
            dys = Split.from_f32(dy)
        def param_grads():
            if bname:      # from the split copy: dy itself may be the residual-stream gradient, which LayerNorm backward updates in place
                ops.colsum_split_(dys, cout, self.grads[bname])
            self.wgrad(x_saved, dys, wname, kind, cin, cout, taps)

        # the parameter gradients are leaves of the backward graph: they run beside the dx chain (the critical path)
        self.on_side(param_grads, dys, x_saved)
        dx = None
        if need_dx:
            dx, _ = ops.conv_gemm(dys, self.w_bwd(wname, kind), n=cin, k=cout, taps=taps)
        return dx

    def on_side(self, fn, *keep):
        """Run fn() on the side stream, after everything issued so far on the current stream (fork); join_side() is the join.
        The small-batch step is launch / latency bound (~750 kernels of 5-30 us on a few SMs each): the weight-gradient
        transposes, split-K GEMMs and bias sums overlap the activation-gradient chain - in the captured graph they become
        parallel branches.  `keep`: tensors fn reads that were allocated on the current stream - held until the join so that the
        caching allocator (also at capture time) cannot hand their memory to a later tensor while the side branch still reads it."""
        if not self.overlap:
            fn()
            return
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        self._side.wait_stream(cur)
        self._keep.extend(keep)
        with torch.cuda.stream(self._side):
            fn()
        self._side_used = True

    def join_side(self):
        if self._side_used:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_used = False
        self._keep.clear()

    def zbuf(self, role, shape):
        """Persistent zero-initialised operand planes of the current batch shape (training/wgrad.py: ZeroPlanes)."""
        return self._zp.get(role, shape, self.dev)

    def wgrad(self, x, dys, wname, kind, cin, cout, taps):
        """dW = X^T dY over the flattened (batch, time) axis, split-K (training/wgrad.py)."""
        B, T = x.hi.shape[0], x.hi.shape[1]
        dev = x.hi.device
        Tp, S, ks, KKp = wgrad.plan(B, T, cout, cin)
        dyt = self.zbuf(("dyt", B, T), (cout, KKp))
        ops.transpose_planes(dys, z=B, rows=T, src_zstride=T * dys.hi.shape[2], ld_src=dys.hi.shape[2], c0=0, cols=cout, shift=0,
                             r_out=T, dst=dyt, dst_zstride=Tp, ld_dst=KKp)
        pad = (taps - 1) // 2
        tmp = torch.empty(taps, cout, cin, dtype=torch.float32, device=dev) if kind == "conv" else None
        for tap in range(taps):
            xt = self.zbuf(("xt", B, T), (cin, KKp))
            ops.transpose_planes(x, z=B, rows=T, src_zstride=x.hi.stride(0), ld_src=x.hi.stride(1), c0=0, cols=cin, shift=tap - pad,
                                 r_out=T, dst=xt, dst_zstride=Tp, ld_dst=KKp)
            if kind == "lin":    # Paddle Linear weight [in, out]
                wgrad.nt_splitk(xt, dyt, cin, cout, S, ks, KKp, out=self.grads[wname])
            else:                # Conv1D weight [out, in, k]
                wgrad.nt_splitk(dyt, xt, cout, cin, S, ks, KKp, out=tmp[tap])
        if kind == "conv":
            self.grads[wname].copy_(tmp.permute(1, 2, 0))

    # ------------------------------------------------------------------------------------------------------------
    # FFT-block stack (Encoder.forward after the embedding) with saved context
    # ------------------------------------------------------------------------------------------------------------
    def stack_fwd(self, x, pre, n_layers, key_lens):
        m = self.m
        sid = 0 if pre == "encoder." else 1
        tag = "enc" if sid == 0 else "dec"
        r_layer, r_attn = self.rates[f"transformer_{tag}_dropout_rate"], self.rates[f"transformer_{tag}_attn_dropout_rate"]
        B, T, A = x.shape
        H, dk = m.aheads, A // m.aheads
        Tp = _ceil64(T)
        ctxs = []
        kind = "lin" if m._linear_ffn else "conv"
        for i in range(n_layers):
            q = f"{pre}encoders.{i}."
            c = dict(x0=x)
            _, c["h1"] = ops.layer_norm(x, self.P(q + "norm1.weight"), self.P(q + "norm1.bias"))
            wqkv = self._pack(("f", q + "qkv"), lambda: pack_dev(torch.cat([self.P(q + "self_attn.linear_q.weight"), self.P(q + "self_attn.linear_k.weight"),
                                                                          self.P(q + "self_attn.linear_v.weight")], dim=1).t().contiguous()))
            bqkv = torch.cat([self.P(q + "self_attn.linear_q.bias"), self.P(q + "self_attn.linear_k.bias"), self.P(q + "self_attn.linear_v.bias")])
            _, qkv = ops.conv_gemm(c["h1"], wqkv, n=3 * A, k=A, bias=bqkv, out_f32=False, out_split=True)
            c["qkv"] = qkv
            ld = 3 * A
            s_buf = torch.empty(B * H, T, Tp, dtype=torch.float32, device=x.device)
            q_spec = dict(rows=T, cols=ld, ld=ld, batch_stride=T * ld, batches=B, bmul=1, hmul=0, col0=0, colh=dk)
            k_spec = dict(rows=T, cols=ld, ld=ld, batch_stride=T * ld, batches=B, bmul=1, hmul=0, col0=A, colh=dk)
            ops.batched_matmul_nt(qkv, qkv, batch=B, heads=H, m=T, n=T, k=dk, a_spec=q_spec, b_spec=k_spec, scale=1.0 / math.sqrt(dk),
                                  y_f32=s_buf, y_batch_stride=H * T * Tp, y_head_stride=T * Tp, y_ld=Tp)
            c["p"] = ops.masked_softmax(s_buf, key_lens, B, H, T, T)
            c["pd"] = self.drop(c["p"], r_attn, self.site(sid, i, 1), out_f32=False, out_split=True)[1] if r_attn > 0 else c["p"]
            vt = ops.transpose_heads(qkv, col0=2 * A, dk=dk, heads=H, ld_dst=Tp)
            ctx = Split.empty((B, T, A), x.device)
            p_spec = dict(rows=T, cols=Tp, ld=Tp, batch_stride=T * Tp, batches=B * H, bmul=H, hmul=1, col0=0, colh=0)
            v_spec = dict(rows=dk, cols=Tp, ld=Tp, batch_stride=dk * Tp, batches=B * H, bmul=H, hmul=1, col0=0, colh=0)
            ops.batched_matmul_nt(c["pd"], vt, batch=B, heads=H, m=T, n=dk, k=Tp, a_spec=p_spec, b_spec=v_spec, y_split=ctx,
                                  y_batch_stride=T * A, y_head_stride=dk, y_ld=A)
            c["ctx"] = ctx
            if r_layer > 0:      # x1 = x + dropout(attention): the residual add cannot ride in the GEMM epilogue any more
                a_out, _ = self.layer_fwd(ctx, q + "self_attn.linear_out.weight", q + "self_attn.linear_out.bias", "lin")
                self.drop(a_out, r_layer, self.site(sid, i, 2), inplace=True)
                ops.axpy_(1.0, x, a_out)
                x1 = a_out
            else:
                x1, _ = self.layer_fwd(ctx, q + "self_attn.linear_out.weight", q + "self_attn.linear_out.bias", "lin", residual=x)
            c["x1"] = x1
            _, c["h2"] = ops.layer_norm(x1, self.P(q + "norm2.weight"), self.P(q + "norm2.bias"))
            _, c["u"] = self.layer_fwd(c["h2"], q + "feed_forward.w_1.weight", q + "feed_forward.w_1.bias", kind, act="relu", f32=False, split=True)
            c["ud"] = self.drop(c["u"], r_layer, self.site(sid, i, 3), out_f32=False, out_split=True)[1] if r_layer > 0 else c["u"]
            if r_layer > 0:
                f_out, _ = self.layer_fwd(c["ud"], q + "feed_forward.w_2.weight", q + "feed_forward.w_2.bias", kind)
                self.drop(f_out, r_layer, self.site(sid, i, 4), inplace=True)
                ops.axpy_(1.0, x1, f_out)
                x = f_out
            else:
                x, _ = self.layer_fwd(c["u"], q + "feed_forward.w_2.weight", q + "feed_forward.w_2.bias", kind, residual=x1)
            ctxs.append(c)
        y, ys = ops.layer_norm(x, self.P(pre + "after_norm.weight"), self.P(pre + "after_norm.bias"), want_f32=True, want_split=True)
        return y, ys, dict(layers=ctxs, x_last=x, pre=pre, n=n_layers, sid=sid, r_layer=r_layer, r_attn=r_attn)

    def stack_bwd(self, dy, S):
        """dy: gradient w.r.t. the after_norm output (fp32).  Returns the gradient w.r.t. the stack input."""
        m = self.m
        pre = S["pre"]
        B, T, A = dy.shape
        H, dk = m.aheads, A // m.aheads
        Tp = _ceil64(T)
        dev = dy.device
        kind = "lin" if m._linear_ffn else "conv"
        dx = torch.empty_like(dy)
        ops.layer_norm_bwd(S["x_last"], self.P(pre + "after_norm.weight"), dy, dx, False, self.grads[pre + "after_norm.weight"],
                           self.grads[pre + "after_norm.bias"])
        sid, r_layer, r_attn = S["sid"], S["r_layer"], S["r_attn"]
        for i in reversed(range(S["n"])):
            q = f"{pre}encoders.{i}."
            c = S["layers"][i]
            # x2 = x1 + drop(conv2(drop(relu(conv1(LN2(x1))))))
            dsub = self.drop(dx, r_layer, self.site(sid, i, 4))[0] if r_layer > 0 else dx
            du = self.layer_bwd(dsub, c["ud"], q + "feed_forward.w_2.weight", q + "feed_forward.w_2.bias", kind)
            if r_layer > 0:
                self.drop(du, r_layer, self.site(sid, i, 3), inplace=True)
            du_f, _ = ops.relu_bwd(du, c["u"], want_f32=True)
            dh2 = self.layer_bwd(du_f, c["h2"], q + "feed_forward.w_1.weight", q + "feed_forward.w_1.bias", kind)
            ops.layer_norm_bwd(c["x1"], self.P(q + "norm2.weight"), dh2, dx, True, self.grads[q + "norm2.weight"], self.grads[q + "norm2.bias"])
            # x1 = x0 + drop(out_proj(attention(LN1(x0))))
            dsub = self.drop(dx, r_layer, self.site(sid, i, 2))[0] if r_layer > 0 else dx
            dctx = self.layer_bwd(dsub, c["ctx"], q + "self_attn.linear_out.weight", q + "self_attn.linear_out.bias", "lin")
            dctx_s = Split.from_f32(dctx)
            qkv, p = c["qkv"], c["p"]
            ld = 3 * A
            dqkv = torch.zeros(B, T, ld, dtype=torch.float32, device=dev)
            o_spec = dict(rows=T, cols=A, ld=A, batch_stride=T * A, batches=B, bmul=1, hmul=0, col0=0, colh=dk)
            v_spec = dict(rows=T, cols=ld, ld=ld, batch_stride=T * ld, batches=B, bmul=1, hmul=0, col0=2 * A, colh=dk)
            dp = torch.zeros(B * H, T, Tp, dtype=torch.float32, device=dev)
            ops.batched_matmul_nt(dctx_s, qkv, batch=B, heads=H, m=T, n=T, k=dk, a_spec=o_spec, b_spec=v_spec, y_f32=dp,
                                  y_batch_stride=H * T * Tp, y_head_stride=T * Tp, y_ld=Tp)                       # d(drop(P)) = dO V^T
            if r_attn > 0:
                self.drop(dp, r_attn, self.site(sid, i, 1), inplace=True)                                         # -> dP
            ds = ops.softmax_bwd(p, dp, T, 1.0 / math.sqrt(dk))                                                  # includes the 1/sqrt(dk)
            z_spec = dict(rows=T, cols=Tp, ld=Tp, batch_stride=T * Tp, batches=B * H, bmul=H, hmul=1, col0=0, colh=0)
            d_spec = dict(rows=dk, cols=Tp, ld=Tp, batch_stride=dk * Tp, batches=B * H, bmul=H, hmul=1, col0=0, colh=0)

            def t_sq(src):       # (B*H, T, Tp) -> transposed (B*H, T, Tp)
                dst = self.zbuf("tsq", (B * H, T, Tp))
                ops.transpose_planes(src, z=B * H, rows=T, src_zstride=T * Tp, ld_src=Tp, c0=0, cols=T, shift=0, r_out=T, dst=dst,
                                     dst_zstride=T * Tp, ld_dst=Tp)
                return dst

            def t_heads(src, ld_src, col0):   # (B, T, ld_src)[.., col0 + h*dk + d] -> (B*H, dk, Tp)
                dst = self.zbuf(("th", T), (B, H, dk, Tp))
                for h in range(H):
                    ops.transpose_planes(src, z=B, rows=T, src_zstride=T * ld_src, ld_src=ld_src, c0=col0 + h * dk, cols=dk, shift=0,
                                         r_out=T, dst=Split(dst.hi[:, h], dst.lo[:, h]), dst_zstride=H * dk * Tp, ld_dst=Tp)
                return dst

            pt, dot = t_sq(c["pd"]), t_heads(dctx_s, A, 0)                                                        # dV uses the dropped P
            ops.batched_matmul_nt(pt, dot, batch=B, heads=H, m=T, n=dk, k=Tp, a_spec=z_spec, b_spec=d_spec, y_f32=dqkv[:, :, 2 * A:],
                                  y_batch_stride=T * ld, y_head_stride=dk, y_ld=ld)                               # dV = P^T dO
            kt = t_heads(qkv, ld, A)
            ops.batched_matmul_nt(ds, kt, batch=B, heads=H, m=T, n=dk, k=Tp, a_spec=z_spec, b_spec=d_spec, y_f32=dqkv,
                                  y_batch_stride=T * ld, y_head_stride=dk, y_ld=ld)                               # dQ = dS K
            dst_, qt = t_sq(ds), t_heads(qkv, ld, 0)
            ops.batched_matmul_nt(dst_, qt, batch=B, heads=H, m=T, n=dk, k=Tp, a_spec=z_spec, b_spec=d_spec, y_f32=dqkv[:, :, A:],
                                  y_batch_stride=T * ld, y_head_stride=dk, y_ld=ld)                               # dK = dS^T Q
            # fused QKV projection: h1 [A] -> [3A]
            dqs = Split.from_f32(dqkv)
            wq_b = self._pack(("b", q + "qkv"), lambda: pack_dev(torch.cat([self.P(q + "self_attn.linear_q.weight"), self.P(q + "self_attn.linear_k.weight"),
                                                                           self.P(q + "self_attn.linear_v.weight")], dim=1).contiguous()))

            def qkv_param_grads(q=q, dqkv=dqkv, dqs=dqs, h1=c["h1"]):
                bsum = torch.zeros(ld, dtype=torch.float32, device=dev)
                ops.colsum_(dqkv.reshape(B * T, ld), bsum)
                for j, nm in enumerate(("linear_q", "linear_k", "linear_v")):
                    self.grads[q + "self_attn." + nm + ".bias"].copy_(bsum[j * A:(j + 1) * A])
                Tq, Sq, ksq, KKq = wgrad.plan(B, T, A, ld)
                xt = self.zbuf(("xt", B, T), (A, KKq))
                ops.transpose_planes(h1, z=B, rows=T, src_zstride=T * A, ld_src=A, c0=0, cols=A, shift=0, r_out=T, dst=xt, dst_zstride=Tq, ld_dst=KKq)
                dyt = self.zbuf(("dyt", B, T), (ld, KKq))
                ops.transpose_planes(dqs, z=B, rows=T, src_zstride=T * ld, ld_src=ld, c0=0, cols=ld, shift=0, r_out=T, dst=dyt, dst_zstride=Tq,
                                     ld_dst=KKq)
                gw = wgrad.nt_splitk(xt, dyt, A, ld, Sq, ksq, KKq)
                for j, nm in enumerate(("linear_q", "linear_k", "linear_v")):
                    self.grads[q + "self_attn." + nm + ".weight"].copy_(gw[:, j * A:(j + 1) * A])

            self.on_side(qkv_param_grads, dqkv, dqs, c["h1"])
            dh1, _ = ops.conv_gemm(dqs, wq_b, n=A, k=ld)
            ops.layer_norm_bwd(c["x0"], self.P(q + "norm1.weight"), dh1, dx, True, self.grads[q + "norm1.weight"], self.grads[q + "norm1.bias"])
        return dx

    # ------------------------------------------------------------------------------------------------------------
    # predictors
    # ------------------------------------------------------------------------------------------------------------
    def pred_fwd(self, pre, n_layers, hs_split):
        sid, rate = {"pitch_predictor.": (2, self.rates["pitch_predictor_dropout"]), "energy_predictor.": (3, self.rates["energy_predictor_dropout"]),
                     "duration_predictor.": (4, self.rates["duration_predictor_dropout_rate"])}[pre]
        saved, h = [], hs_split
        for i in range(n_layers):
            y, ys = self.layer_fwd(h, f"{pre}conv.{i}.0.weight", f"{pre}conv.{i}.0.bias", "conv", act="relu", f32=True, split=True)
            _, hn = ops.layer_norm(y, self.P(f"{pre}conv.{i}.2.weight"), self.P(f"{pre}conv.{i}.2.bias"))
            if rate > 0:
                hn = self.drop(hn, rate, self.site(sid, i, 5), out_f32=False, out_split=True)[1]
            saved.append(dict(x=h, y=y, ys=ys))
            h = hn
        out, _ = self.layer_fwd(h, pre + "linear.weight", pre + "linear.bias", "lin")
        return out, dict(layers=saved, h_last=h, pre=pre, sid=sid, rate=rate)

    def pred_bwd(self, dout, S, need_dx=True):
        pre = S["pre"]
        g = self.layer_bwd(dout, S["h_last"], pre + "linear.weight", pre + "linear.bias", "lin")
        n = len(S["layers"])
        for i in reversed(range(n)):
            c = S["layers"][i]
            if S["rate"] > 0:
                self.drop(g, S["rate"], self.site(S["sid"], i, 5), inplace=True)
            dy = torch.empty_like(g)
            ops.layer_norm_bwd(c["y"], self.P(f"{pre}conv.{i}.2.weight"), g, dy, False, self.grads[f"{pre}conv.{i}.2.weight"],
                               self.grads[f"{pre}conv.{i}.2.bias"])
            dpre, _ = ops.relu_bwd(dy, c["ys"], want_f32=True)
            g = self.layer_bwd(dpre, c["x"], f"{pre}conv.{i}.0.weight", f"{pre}conv.{i}.0.bias", "conv", need_dx=(i > 0 or need_dx))
        return g

    # ------------------------------------------------------------------------------------------------------------
    # one training step
    # ------------------------------------------------------------------------------------------------------------
    def forward_backward(self, batch):
        m = self.m
        L = _lib.lib()
        st = _stream()
        dev = m.device
        self._packs = {}
        self._zp.begin(self._batch_key(batch))
        self.gflat.zero_()
        text = batch["text"].to(dev, torch.int64).contiguous()
        B, T = text.shape
        ilens = _i32(batch["text_lengths"].to(dev))
        olens = _i32(batch["speech_lengths"].to(dev))
        ds = batch["durations"].to(dev, torch.int64).contiguous()
        ps = batch["pitch"].to(dev, torch.float32).reshape(B, T).contiguous()
        es = batch["energy"].to(dev, torch.float32).reshape(B, T).contiguous()
        ys = batch["speech"].to(dev, torch.float32).contiguous()
        A, odim = m.adim, m.odim
        # ---- forward (train mode) ----
        R = self.rates
        x = ops.embed_pe(text, self.P("encoder.embed.0.weight"), None, self.P("encoder.embed.1.alpha"), None, m.padding_idx)
        if R["transformer_enc_positional_dropout_rate"] > 0:
            self.drop(x, R["transformer_enc_positional_dropout_rate"], self.site(0, 0, 0), inplace=True)
        hs, hs_split, S_enc = self.stack_fwd(x, "encoder.", m.elayers, ilens)
        p_raw, S_p = self.pred_fwd("pitch_predictor.", m.cfg["pitch"][0], hs_split)
        e_raw, S_e = self.pred_fwd("energy_predictor.", m.cfg["energy"][0], hs_split)
        d_raw, S_d = self.pred_fwd("duration_predictor.", m.cfg["dur"][0], hs_split)
        p_outs = ops.mask_rows_(p_raw.reshape(B, T).clone(), ilens)
        e_outs = ops.mask_rows_(e_raw.reshape(B, T).clone(), ilens)
        d_outs = ops.mask_rows_(d_raw.reshape(B, T).clone(), ilens)
        pe_w, ee_w = self.P("pitch_embed.0.weight"), self.P("energy_embed.0.weight")
        r_pe, r_ee = R["pitch_embed_dropout"], R["energy_embed_dropout"]
        if r_pe > 0 or r_ee > 0:
            # Sequential(Conv1D, Dropout) (fastspeech2.py:220-247): the two embeddings one at a time (the fused kernel with the
            # other embedding's weights zeroed), each through its own mask, then hs + e_embs + p_embs
            zw_p, zw_e = torch.zeros_like(pe_w.reshape(A, -1)), torch.zeros_like(ee_w.reshape(A, -1))
            zb, z0 = torch.zeros(A, device=dev), torch.zeros_like(hs)
            p_emb = ops.variance_embed_add(z0, ps, es, pe_w.reshape(A, -1), self.P("pitch_embed.0.bias"), zw_e, zb, None)
            e_emb = ops.variance_embed_add(z0, ps, es, zw_p, zb, ee_w.reshape(A, -1), self.P("energy_embed.0.bias"), None)
            if r_pe > 0:
                self.drop(p_emb, r_pe, self.site(6, 0, 7), inplace=True)
            if r_ee > 0:
                self.drop(e_emb, r_ee, self.site(6, 0, 8), inplace=True)
            hs2 = hs.clone()
            ops.axpy_(1.0, e_emb, hs2)
            ops.axpy_(1.0, p_emb, hs2)
        else:
            hs2 = ops.variance_embed_add(hs, ps, es, pe_w.reshape(A, -1), self.P("pitch_embed.0.bias"), ee_w.reshape(A, -1),
                                         self.P("energy_embed.0.bias"), None)
        t_dec = ys.shape[1]
        hs_lr, _ = ops.length_regulate(hs2, ds, t_dec)
        xd = ops.embed_pe(None, None, hs_lr, self.P("decoder.embed.0.alpha"), None)
        if R["transformer_dec_positional_dropout_rate"] > 0:
            self.drop(xd, R["transformer_dec_positional_dropout_rate"], self.site(1, 0, 0), inplace=True)
        zs, zs_split, S_dec = self.stack_fwd(xd, "decoder.", m.dlayers, olens)
        before, before_split = self.layer_fwd(zs_split, "feat_out.weight", "feat_out.bias", "lin", f32=True, split=True)
        post, h = [], before_split
        rows = B * t_dec
        for i in range(m.postnet_layers):
            last = i == m.postnet_layers - 1
            cw = f"postnet.postnet.{i}.0.weight"
            q = f"postnet.postnet.{i}.1."
            conv_out, _ = self.layer_fwd(h, cw, None, "conv")
            cdim = conv_out.shape[-1]
            y = torch.empty_like(conv_out)
            ysplit = Split.empty(tuple(conv_out.shape), dev) if not last else None
            mean = torch.empty(cdim, device=dev)
            rstd = torch.empty(cdim, device=dev)
            _lib.check(L.pk_batch_norm_train(_ptr(conv_out), rows, cdim, _ptr(self.P(q + "weight")), _ptr(self.P(q + "bias")), 1e-5,
                                             0 if last else 2, 0.9, _ptr(m._params[q + "_mean"]), _ptr(m._params[q + "_variance"]),
                                             _ptr(self.sums), _ptr(y), _ptr(ysplit.hi) if ysplit else None,
                                             _ptr(ysplit.lo) if ysplit else None, _ptr(mean), _ptr(rstd), st), "pk_batch_norm_train")
            yd = y
            if R["postnet_dropout_rate"] > 0:                # Dropout closes every postnet layer (tacotron2/decoder.py:144-180)
                yd, ysplit = self.drop(y, R["postnet_dropout_rate"], self.site(5, i, 6), out_f32=True, out_split=not last)
            post.append(dict(x=h, conv=conv_out, y=y, yd=yd, mean=mean, rstd=rstd))
            h = ysplit
        after = before.clone()
        ops.axpy_(1.0, post[-1]["yd"], after)
        # ---- loss and its gradient ----
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        ws = torch.empty(12, dtype=torch.float32, device=dev)
        _lib.check(L.pk_fs2_loss(_ptr(before), _ptr(after), _ptr(ys), _ptr(olens), t_dec, odim, _ptr(d_outs), _ptr(ds), _ptr(p_outs), _ptr(ps),
                                 _ptr(e_outs), _ptr(es), _ptr(ilens), T, B, _ptr(ws), _ptr(losses), st), "pk_fs2_loss")
        g_before, g_after = torch.empty_like(before), torch.empty_like(after)
        g_d, g_p, g_e = torch.empty(B, T, 1, device=dev), torch.empty(B, T, 1, device=dev), torch.empty(B, T, 1, device=dev)
        _lib.check(L.pk_fs2_loss_bwd(_ptr(before), _ptr(after), _ptr(ys), _ptr(olens), t_dec, odim, _ptr(d_outs), _ptr(ds), _ptr(p_outs),
                                     _ptr(ps), _ptr(e_outs), _ptr(es), _ptr(ilens), T, B, _ptr(g_before), _ptr(g_after), _ptr(g_d), _ptr(g_p),
                                     _ptr(g_e), st), "pk_fs2_loss_bwd")
        # ---- backward ----
        g = g_after                                            # after = before + postnet(before)
        for i in reversed(range(m.postnet_layers)):
            last = i == m.postnet_layers - 1
            c = post[i]
            q = f"postnet.postnet.{i}.1."
            cdim = c["conv"].shape[-1]
            dconv = torch.empty_like(c["conv"])
            if R["postnet_dropout_rate"] > 0:
                g = self.drop(g, R["postnet_dropout_rate"], self.site(5, i, 6))[0]
            _lib.check(L.pk_batch_norm_bwd(_ptr(c["conv"]), _ptr(g), _ptr(c["y"]), _ptr(c["mean"]), _ptr(c["rstd"]), _ptr(self.P(q + "weight")),
                                           0 if last else 2, rows, cdim, _ptr(self.sums), _ptr(dconv), st), "pk_batch_norm_bwd")
            self.grads[q + "bias"].copy_(self.sums[:cdim])
            self.grads[q + "weight"].copy_(self.sums[cdim:2 * cdim])
            g = self.layer_bwd(dconv, c["x"], f"postnet.postnet.{i}.0.weight", None, "conv")
        ops.axpy_(1.0, g_after, g)                             # residual path of `after`
        ops.axpy_(1.0, g_before, g)                            # direct L1 on `before`
        dzs = self.layer_bwd(g, zs_split, "feat_out.weight", "feat_out.bias", "lin")
        dxd = self.stack_bwd(dzs, S_dec)
        if R["transformer_dec_positional_dropout_rate"] > 0:
            self.drop(dxd, R["transformer_dec_positional_dropout_rate"], self.site(1, 0, 0), inplace=True)
        _lib.check(L.pk_embed_pe_bwd(None, _ptr(dxd), 0, 0, B, t_dec, A, None, _ptr(self.grads["decoder.embed.0.alpha"]), st), "pk_embed_pe_bwd")
        dhs = torch.empty(B, T, A, dtype=torch.float32, device=dev)
        _lib.check(L.pk_length_regulate_bwd(_ptr(dxd), _ptr(ds), B, T, A, t_dec, _ptr(dhs), st), "pk_length_regulate_bwd")
        kp, ke = pe_w.shape[-1], ee_w.shape[-1]
        d_pe = self.drop(dhs, r_pe, self.site(6, 0, 7))[0] if r_pe > 0 else dhs
        d_ee = self.drop(dhs, r_ee, self.site(6, 0, 8))[0] if r_ee > 0 else dhs
        _lib.check(L.pk_scalar_conv_wgrad(_ptr(d_pe), _ptr(ps), B, T, A, kp, _ptr(self.grads["pitch_embed.0.weight"]),
                                          _ptr(self.grads["pitch_embed.0.bias"]), st), "pk_scalar_conv_wgrad")
        _lib.check(L.pk_scalar_conv_wgrad(_ptr(d_ee), _ptr(es), B, T, A, ke, _ptr(self.grads["energy_embed.0.weight"]),
                                          _ptr(self.grads["energy_embed.0.bias"]), st), "pk_scalar_conv_wgrad")
        gd = self.pred_bwd(g_d, S_d)
        ops.axpy_(1.0, gd, dhs)
        ge = self.pred_bwd(g_e, S_e, need_dx=not self.sg_energy)
        if not self.sg_energy:
            ops.axpy_(1.0, ge, dhs)
        gp = self.pred_bwd(g_p, S_p, need_dx=not self.sg_pitch)
        if not self.sg_pitch:
            ops.axpy_(1.0, gp, dhs)
        dx = self.stack_bwd(dhs, S_enc)
        if R["transformer_enc_positional_dropout_rate"] > 0:
            self.drop(dx, R["transformer_enc_positional_dropout_rate"], self.site(0, 0, 0), inplace=True)
        _lib.check(L.pk_embed_pe_bwd(_ptr(text), _ptr(dx), m.idim, m.padding_idx, B, T, A, _ptr(self.grads["encoder.embed.0.weight"]),
                                     _ptr(self.grads["encoder.embed.1.alpha"]), st), "pk_embed_pe_bwd")
        self.join_side()
        return losses

    _BATCH_ORDER = ("text", "text_lengths", "speech", "speech_lengths", "durations", "pitch", "energy")

    @classmethod
    def _batch_key(cls, batch):
        return tuple(tuple(batch[k].shape) for k in cls._BATCH_ORDER)

    def _forward_backward_graphed(self, batch):
        dev = self.m.device
        order = self._BATCH_ORDER
        dtypes = (torch.int64, torch.int64, torch.float32, torch.int64, torch.int64, torch.float32, torch.float32)
        tensors = [batch[k].to(dev, dt).contiguous() for k, dt in zip(order, dtypes)]
        key = tuple(tuple(t.shape) for t in tensors)
        self._zp.touch(key)                      # a replay does not pass through forward_backward: keep the LRU order honest
        fn = lambda *ts: self.forward_backward(dict(zip(order, ts)))
        return self._fb_graphs.run(key, fn, tensors).clone()

    # ------------------------------------------------------------------------------------------------------------
    # snapshot / resume (reference: StandardUpdater.state_dict / set_state_dict, training/updaters/standard_updater.py;
    # Snapshot extension writes it with paddle.save as snapshot_iter_<n>.pdz and train.py resumes by constructing the
    # updater first and loading afterwards - which is why Layer.set_state_dict copies IN PLACE into the flat buffer)
    # ------------------------------------------------------------------------------------------------------------
    def state_dict(self, epoch=0):
        """{"main_params", "main_optimizer", "epoch", "iteration"}: Adam moments per parameter under Paddle's accumulator
        suffixes (`<name>_moment1_0`, `<name>_moment2_0`; Paddle prefixes them with its internal tensor names, which do not
        exist here, so the structured names are used) plus the step count the bias correction needs."""
        opt = {}
        for k, o, n in zip(self.buffers.names, self.buffers.offsets, self.buffers.sizes):
            shape = self.m._params[k].shape
            opt[k + "_moment1_0"] = self.adam_m[o:o + n].view(shape).clone()
            opt[k + "_moment2_0"] = self.adam_v[o:o + n].view(shape).clone()
        opt["step_count"] = self.step_count
        opt["LR_Scheduler"] = {"last_lr": self.lr}
        return {"main_params": self.m.state_dict(), "main_optimizer": opt, "epoch": int(epoch), "iteration": int(self.step_count)}

    def set_state_dict(self, state):
        self.m.set_state_dict(state["main_params"])                  # in place: the parameters stay views of self.flat
        opt = state.get("main_optimizer", {})
        for k, o, n in zip(self.buffers.names, self.buffers.offsets, self.buffers.sizes):
            for suffix, buf in (("_moment1_0", self.adam_m), ("_moment2_0", self.adam_v)):
                if k + suffix in opt:
                    buf[o:o + n].copy_(torch.as_tensor(opt[k + suffix]).reshape(-1).to(buf.device, buf.dtype))
        self.step_count = int(opt.get("step_count", state.get("iteration", self.step_count)))
        self.step_dev.fill_(self.step_count)
        self._packs = {}

    def step(self, batch):
        """One update: returns the four loss values (device tensor: l1, duration, pitch, energy)."""
        losses = self._forward_backward_graphed(batch) if self.use_graphs else self.forward_backward(batch)
        if self.world > 1:
            self.buffers.all_reduce_grads(self.group)                                # the one exchange step of the path
        self.step_count += 1
        _lib.check(_lib.lib().pk_adam(_ptr(self.flat), _ptr(self.gflat), _ptr(self.adam_m), _ptr(self.adam_v), self.flat.numel(),
                                      self.lr, self.b1, self.b2, self.eps, self.step_count, 1.0 / self.world, _stream()), "pk_adam")
        self.step_dev += 1
        self.m._packed = None
        return losses
