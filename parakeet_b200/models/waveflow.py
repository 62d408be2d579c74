"""ConditionalWaveFlow inference on B200 - host side (reference parakeet/models/waveflow.py:714-909).

Same constructor as the reference (`upsample_factors, n_flows, n_layers, n_group, channels, n_mels, kernel_size`), same
state-dict keys (`encoder.{i}.{weight_g,weight_v,bias}`, `decoder.{f}.input_proj.*`,
`decoder.{f}.resnet.{l}.{conv,condition_proj,out_proj}.*`, `decoder.{f}.output_proj.{weight,bias}`), `infer(mel)` and
`predict(mel)`.  All arithmetic is in libparakeet_b200.so; fold / permute / row slicing are torch views and gathers.

Supported this round: kernel_size (3, 3) and n_group in {8, 16} (height dilation 1 -> a 3-row causal buffer), which
covers the shipped config (examples/waveflow/config.py).  Training (`forward`, WaveFlowLoss) is not implemented.
"""
import ctypes as C_
import os

import numpy as np
import torch

from .. import _lib, ops
from ..layer import Layer
from ..ops import Split, _ptr, _stream


def _planes(m, dev):
    """fp32 matrix -> split-bf16 planes (2, rows, cols) in ONE allocation (the 4-D TMA maps of the fused kernels need it)."""
    m = m.detach().float().cpu()
    hi = m.to(torch.bfloat16)
    lo = (m - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous().to(dev)


def _fold_wn(params):
    out = {}
    for k, v in params.items():
        if k.endswith("weight_g"):
            continue
        if k.endswith("weight_v"):
            g = params[k[:-1] + "g"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1)
            out[k[:-2]] = v * (g / norm).reshape([-1] + [1] * (v.dim() - 1))
        else:
            out[k] = v
    return out


class ConditionalWaveFlow(Layer):
    def __init__(self, upsample_factors, n_flows, n_layers, n_group, channels, n_mels, kernel_size=(3, 3), device=None, seed=0):
        super().__init__(device)
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        if tuple(kernel_size) != (3, 3) or n_group not in (8, 16):
            raise NotImplementedError("this round supports kernel_size (3,3) and n_group 8 / 16 (height dilation 1)")
        if n_group % 2 or n_flows % 2:
            raise ValueError("number of flows and number of group must be even")
        if channels % 64:
            raise NotImplementedError("channels must be a multiple of 64")
        self.upsample_factors = list(upsample_factors)
        self.n_flows, self.n_layers, self.n_group, self.channels, self.n_mels = n_flows, n_layers, n_group, channels, n_mels
        g = torch.Generator().manual_seed(seed)

        def u(*shape, std):
            return (torch.rand(*shape, generator=g) * 2 - 1) * std

        def wn(name, w):
            self._register(name + ".weight_g", w.reshape(w.shape[0], -1).norm(dim=1))
            self._register(name + ".weight_v", w)

        import math
        for i, f in enumerate(self.upsample_factors):
            std = math.sqrt(1 / (3 * 2 * f))
            wn(f"encoder.{i}", u(1, 1, 3, 2 * f, std=std))
            self._register(f"encoder.{i}.bias", u(1, std=std))
        C = channels
        for fl in range(n_flows):
            pre = f"decoder.{fl}."
            wn(pre + "input_proj", u(C, 1, 1, 1, std=1.0))
            self._register(pre + "input_proj.bias", u(C, std=1.0))
            for l in range(n_layers):
                q = f"{pre}resnet.{l}."
                std = math.sqrt(1 / (C * 9))
                wn(q + "conv", u(2 * C, C, 3, 3, std=std))
                self._register(q + "conv.bias", u(2 * C, std=std))
                std = math.sqrt(1 / n_mels)
                wn(q + "condition_proj", u(2 * C, n_mels, 1, 1, std=std))
                self._register(q + "condition_proj.bias", u(2 * C, std=std))
                std = math.sqrt(1 / C)
                wn(q + "out_proj", u(2 * C, C, 1, 1, std=std))
                self._register(q + "out_proj.bias", u(2 * C, std=std))
            self._register(pre + "output_proj.weight", torch.zeros(2, C, 1, 1))   # reference: Constant(0.)
            self._register(pre + "output_proj.bias", torch.zeros(2))
        idx = list(range(n_group))
        half = n_group // 2
        self.perms = [idx[::-1] if i < n_flows // 2 else list(reversed(idx[:half])) + list(reversed(idx[half:]))
                      for i in range(n_flows)]                                       # waveflow.py:602-615

    def _pack(self):
        if self._packed is not None:
            return self._packed
        p = {k: v.detach().float().cpu() for k, v in _fold_wn(self._params).items()}
        dev, C = self.device, self.channels
        pk = {"enc": [(p[f"encoder.{i}.weight"].reshape(3, -1).contiguous().to(dev), p[f"encoder.{i}.bias"].to(dev))
                      for i in range(len(self.upsample_factors))], "flows": []}
        for fl in range(self.n_flows):
            pre = f"decoder.{fl}."
            layers = []
            for l in range(self.n_layers):
                q = f"{pre}resnet.{l}."
                w = p[q + "conv.weight"]                                  # [2C, C, kh, kw]
                variants = []
                for v in range(3):                                        # row step i with i % 3 == v: slot s holds kh = (s - i) % 3
                    wk = torch.zeros(2 * C, 3 * C, 3)
                    for s in range(3):
                        wk[:, s * C:(s + 1) * C, :] = w[:, :, (s - v) % 3, :]
                    variants.append(ops.pack_weight(wk, dev))
                fused = None
                if self._eligible():
                    # operands of pk_waveflow_flow / pk_waveflow_layer (include/parakeet_b200.h): channels in blocks of 64; gate rows
                    # a_blk | g_blk per block, out_proj rows skip_blk | res_blk; GEMM1 columns [tap][slot][c] | condition_proj
                    nb = C // 64
                    g_rows = torch.cat([torch.cat([torch.arange(64 * k, 64 * k + 64), torch.arange(C + 64 * k, C + 64 * k + 64)])
                                        for k in range(nb)])
                    o_rows = torch.cat([torch.cat([torch.arange(C + 64 * k, C + 64 * k + 64), torch.arange(64 * k, 64 * k + 64)])
                                        for k in range(nb)])
                    cw = p[q + "condition_proj.weight"][:, :, 0, 0]
                    w1 = []
                    for v in range(3):
                        m = torch.zeros(2 * C, 9 * C + 128)
                        for tap in range(3):
                            for s in range(3):
                                m[:, (3 * tap + s) * C:(3 * tap + s + 1) * C] = w[:, :, (s - v) % 3, tap]
                        m[:, 9 * C:9 * C + self.n_mels] = cw
                        w1.append(_planes(m[g_rows], dev))
                    ow, ob = p[q + "out_proj.weight"][:, :, 0, 0], p[q + "out_proj.bias"]
                    fused = dict(w1=w1, w2=_planes(ow[o_rows], dev),
                                 b1=(p[q + "conv.bias"] + p[q + "condition_proj.bias"])[g_rows].numpy().astype("float32").copy(),
                                 b2=ob[o_rows].numpy().astype("float32").copy())
                layers.append(dict(fused=fused, conv=variants, conv_b=p[q + "conv.bias"].to(dev),
                                   cond=ops.pack_weight(p[q + "condition_proj.weight"][:, :, 0, 0], dev),
                                   cond_b=p[q + "condition_proj.bias"].to(dev),
                                   out=ops.pack_weight(p[q + "out_proj.weight"][:, :, 0, 0], dev), out_b=p[q + "out_proj.bias"].to(dev)))
            # the condition projections of all layers of a flow in one GEMM per row step (they do not depend on the recurrence)
            cond_w = torch.cat([p[f"{pre}resnet.{l}.condition_proj.weight"][:, :, 0, 0] for l in range(self.n_layers)], dim=0)
            cond_b = torch.cat([p[f"{pre}resnet.{l}.condition_proj.bias"] for l in range(self.n_layers)])
            f32 = lambda t: t.detach().float().contiguous().numpy().astype("float32").copy()
            host = dict(in_w=f32(p[pre + "input_proj.weight"].reshape(-1)), in_b=f32(p[pre + "input_proj.bias"]),
                        out_w=f32(p[pre + "output_proj.weight"].reshape(2, C)), out_b=f32(p[pre + "output_proj.bias"]))
            pk["flows"].append(dict(host=host, in_w=p[pre + "input_proj.weight"].reshape(-1).contiguous().to(dev),
                                    in_b=p[pre + "input_proj.bias"].to(dev), layers=layers,
                                    cond_all=ops.pack_weight(cond_w, dev), cond_all_b=cond_b.contiguous().to(dev),
                                    out_w=p[pre + "output_proj.weight"].reshape(2, C).contiguous().to(dev),
                                    out_b=p[pre + "output_proj.bias"].to(dev)))
        pk["perms"] = [torch.tensor(pm, dtype=torch.int64, device=dev) for pm in self.perms]   # device-side gather indices
        self._packed = pk
        return pk

    def _fusable(self):
        """The fused kernels cover 64 < n_mels <= 128 and up to 8 layers per flow: pk_waveflow_flow (one persistent launch per
        flow) for 64 or 128 residual channels, pk_waveflow_layer (one launch per ResidualBlock.add_input, PK_WF_FUSED=layer)
        for 64.  PK_WF_FUSED=0 selects the two-GEMM path (A/B runs; also what other channel counts use)."""
        mode = os.environ.get("PK_WF_FUSED", "1")
        return self._eligible() and mode != "0" and (mode != "layer" or self.channels == 64)

    def _eligible(self):
        return self.channels in (64, 128) and 64 < self.n_mels <= 128 and self.n_mels % 8 == 0 and self.n_layers <= 8

    def _flow_mode(self):
        return self._fusable() and os.environ.get("PK_WF_FUSED", "1") != "layer"

    def _run_flow(self, fw, z, x, cond_s, cmap, bufs, skip, flags, prof, st):
        """Rows 1 .. G-1 of one flow in one launch (row 0 and the ring contents are prepared by the caller)."""
        L = _lib.lib()
        B, G, W = z.shape
        NL = self.n_layers
        lay = [l_["fused"] for l_ in fw["layers"]]
        vpa = lambda ptrs: (C_.c_void_p * len(ptrs))(*ptrs)
        keep = dict(cond_rows=(C_.c_int32 * G)(*cmap), ring_hi=vpa([_ptr(b_.hi) for b_ in bufs]), ring_lo=vpa([_ptr(b_.lo) for b_ in bufs]),
                    w1_hi=vpa([_ptr(f["w1"][v][0]) for f in lay for v in range(3)]),
                    w1_lo=vpa([_ptr(f["w1"][v][1]) for f in lay for v in range(3)]),
                    w2_hi=vpa([_ptr(f["w2"][0]) for f in lay]), w2_lo=vpa([_ptr(f["w2"][1]) for f in lay]),
                    bias1=vpa([f["b1"].ctypes.data for f in lay]), bias2=vpa([f["b2"].ctypes.data for f in lay]))
        a = _lib.WaveflowFlowArgs()
        a.batch, a.width, a.channels, a.n_mels, a.n_layers, a.n_group = B, W, self.channels, self.n_mels, NL, G
        for k, v in keep.items():
            setattr(a, k, C_.cast(v, C_.c_void_p))
        a.cond_hi, a.cond_lo = _ptr(cond_s.hi), _ptr(cond_s.lo)
        h = fw["host"]
        a.in_w, a.in_b, a.out_w, a.out_b = (h[k].ctypes.data for k in ("in_w", "in_b", "out_w", "out_b"))
        a.z, a.x, a.skip, a.flags, a.flags_len = _ptr(z), _ptr(x), _ptr(skip), _ptr(flags), flags.numel()
        if prof is not None:
            a.prof = _ptr(prof)
        _lib.check(L.pk_waveflow_flow(C_.byref(a), st), "pk_waveflow_flow")

    def encode(self, mel, trim_conv_artifact=True):
        """UpsampleNet.forward (:103-132): (B, n_mels, T') -> (B, n_mels, T)."""
        pk = self._pack()
        x = mel.contiguous().float()
        L = _lib.lib()
        for (w, b), f in zip(pk["enc"], self.upsample_factors):
            B, Cm, Tin = x.shape
            y = torch.empty(B, Cm, Tin * f - (f if trim_conv_artifact else 0), device=x.device)
            _lib.check(L.pk_waveflow_upsample(_ptr(x), _ptr(w), _ptr(b), B, Cm, Tin, f, 1 if trim_conv_artifact else 0, 0.4,
                                              _ptr(y), _stream()), "pk_waveflow_upsample")
            x = y
        return x

    def inverse(self, z, condition):
        """WaveFlow.inverse (:674-711): z (B, T), condition (B, n_mels, T) -> audio (B, T')."""
        pk = self._pack()
        L = _lib.lib()
        G, C, NL = self.n_group, self.channels, self.n_layers
        pruned = z.shape[-1] // G * G
        z, condition = z[:, :pruned].float(), condition[:, :, :pruned]
        B = z.shape[0]
        W = pruned // G
        dev = z.device
        z = z.reshape(B, W, G).transpose(1, 2).contiguous()                               # (B, H, W): sample t = w*G + h
        cond = condition.reshape(B, self.n_mels, W, G).permute(0, 3, 2, 1).contiguous()   # (B, H, W, n_mels) channels-last
        cond_s = Split.from_f32(cond)
        cmap = list(range(G))                                                             # cumulative row permutation of the condition
        state = torch.empty(B, W, C, device=dev)
        skip = torch.empty(B, W, C, device=dev)
        h_all = torch.empty(B, W, NL * 2 * C, device=dev)                                 # condition projections of one row step
        zt = Split.empty((B, W, C), dev)
        bufs = [Split.zeros((B, W, 3 * C), dev) for _ in range(NL)]
        st = _stream()
        fused = self._fusable()
        flow_mode = self._flow_mode()
        flags = torch.empty((G - 1) * NL * B * ((W + 255) // 256), dtype=torch.int32, device=dev) if flow_mode else None
        prof = getattr(self, "_prof", None)                                               # debug: device uint64[8] phase counters
        for fi in reversed(range(self.n_flows)):
            perm = self.perms[fi]
            z = z.index_select(1, pk["perms"][fi])                                        # geo.shuffle_dim(z, 2, perm)
            cmap = [cmap[j] for j in perm]
            fw = pk["flows"][fi]
            x = torch.empty_like(z)
            x[:, 0] = z[:, 0]
            for b_ in bufs:
                b_.hi.zero_()
                b_.lo.zero_()
            if flow_mode:
                z = z.contiguous()
                _lib.check(L.pk_waveflow_input_proj(_ptr(x[:, 0]), G * W, _ptr(fw["in_w"]), _ptr(fw["in_b"]), B, W, C,
                                                    _ptr(state), _ptr(bufs[0].hi), _ptr(bufs[0].lo), 3 * C, 0, st),
                           "pk_waveflow_input_proj")
                flags.zero_()
                self._run_flow(fw, z, x, cond_s, cmap, bufs, skip, flags, prof, st)
                z = x
                continue
            for i in range(1, G):
                slot = (i - 1) % 3                                                        # ring slot of the newest row (row i-1)
                _lib.check(L.pk_waveflow_input_proj(_ptr(x[:, i - 1]), G * W, _ptr(fw["in_w"]), _ptr(fw["in_b"]), B, W, C,
                                                    _ptr(state), _ptr(bufs[0].hi), _ptr(bufs[0].lo), 3 * C, slot * C, st),
                           "pk_waveflow_input_proj")
                c_row = Split(cond_s.hi[:, cmap[i]], cond_s.lo[:, cmap[i]])              # (B, W, n_mels) views, batch stride G*W*n_mels
                if fused:
                    for l, lay in enumerate(fw["layers"]):
                        f = lay["fused"]
                        a = _lib.WaveflowLayerArgs()
                        a.batch, a.width, a.channels, a.n_mels, a.dilation, a.slot = B, W, C, self.n_mels, 2 ** l, slot
                        a.buf_hi, a.buf_lo = _ptr(bufs[l].hi), _ptr(bufs[l].lo)
                        a.cond_hi, a.cond_lo, a.cond_batch_stride = _ptr(c_row.hi), _ptr(c_row.lo), G * W * self.n_mels
                        w1 = f["w1"][i % 3]
                        a.w1_hi, a.w1_lo, a.w2_hi, a.w2_lo = _ptr(w1[0]), _ptr(w1[1]), _ptr(f["w2"][0]), _ptr(f["w2"][1])
                        a.bias1, a.bias2 = f["b1"].ctypes.data, f["b2"].ctypes.data
                        if l + 1 < NL:
                            a.next_hi, a.next_lo = _ptr(bufs[l + 1].hi), _ptr(bufs[l + 1].lo)
                        a.skip, a.skip_init = _ptr(skip), 1 if l == 0 else 0
                        if prof is not None:
                            a.prof = _ptr(prof)
                        _lib.check(L.pk_waveflow_layer(C_.byref(a), st), "pk_waveflow_layer")
                    _lib.check(L.pk_waveflow_row_out(_ptr(skip), _ptr(fw["out_w"]), _ptr(fw["out_b"]), _ptr(z[:, i]), G * W, B, W, C,
                                                     _ptr(x[:, i]), G * W, st), "pk_waveflow_row_out")
                    continue
                ops.conv_gemm(c_row, fw["cond_all"], n=NL * 2 * C, k=self.n_mels, bias=fw["cond_all_b"], y_f32=h_all)
                for l, lay in enumerate(fw["layers"]):
                    # dilated conv over the 3-row ring + condition slice -> tanh * sigmoid, fused in the GEMM epilogue
                    ops.conv_gemm(bufs[l], lay["conv"][i % 3], n=2 * C, k=3 * C, taps=3, dil=2 ** l, bias=lay["conv_b"], y_split=zt,
                                  epilogue=dict(mode="gate", channels=C, residual=h_all[:, :, l * 2 * C:(l + 1) * 2 * C]))
                    # out_proj -> (res | skip): state += res, skip += skip, next layer's ring slot <- new state, in the epilogue
                    nxt = bufs[l + 1] if l + 1 < NL else None
                    ops.conv_gemm(zt, lay["out"], n=2 * C, k=C, bias=lay["out_b"],
                                  epilogue=dict(mode="wf_update", channels=C, state=state, skip=skip, skip_init=(l == 0), buf=nxt,
                                                buf_col0=slot * C))
                _lib.check(L.pk_waveflow_row_out(_ptr(skip), _ptr(fw["out_w"]), _ptr(fw["out_b"]), _ptr(z[:, i]), G * W, B, W, C,
                                                 _ptr(x[:, i]), G * W, st), "pk_waveflow_row_out")
            z = x
        return z.transpose(1, 2).reshape(B, -1)

    def infer(self, mel, z=None):
        """reference :784-805; the noise z (B, T_c) may be supplied (parity tests), else torch.randn."""
        if not mel.is_cuda:
            raise _lib.PkError("ConditionalWaveFlow needs CUDA tensors (no CPU fallback)")
        mel = mel.contiguous().float()
        B, _, frames = mel.shape
        t_c = frames
        for f in self.upsample_factors:                                     # each transposed conv trims f samples (:121-126)
            t_c = t_c * f - f
        if z is None:
            z = torch.randn(B, t_c, device=mel.device)
        assert z.shape == (B, t_c), (tuple(z.shape), (B, t_c))
        # one CUDA graph per (B, frames): the 5 040 small launches of the row-by-row inverse replay without host work
        fn = lambda m_, z_: self.inverse(z_, self.encode(m_, trim_conv_artifact=True))
        return self._graphs.run(("infer", B, frames), fn, [mel, z.contiguous().float()]).clone()

    @classmethod
    def from_pretrained(cls, config, checkpoint_path, device=None):
        """reference :827-852: build from a config (attribute or mapping access: config.model.*, config.data.n_mels) and load
        `checkpoint_path + ".pdparams"` (utils/checkpoint.py:load_parameters appends the extension) without PaddlePaddle."""
        import os
        from .. import checkpoint

        def get(node, key):
            return node[key] if isinstance(node, dict) else getattr(node, key)
        mc, dc = get(config, "model"), get(config, "data")
        model = cls(upsample_factors=list(get(mc, "upsample_factors")), n_flows=get(mc, "n_flows"), n_layers=get(mc, "n_layers"),
                    n_group=get(mc, "n_group"), channels=get(mc, "channels"), n_mels=get(dc, "n_mels"),
                    kernel_size=tuple(get(mc, "kernel_size")), device=device)
        path = str(checkpoint_path)
        if not os.path.exists(path) and os.path.exists(path + ".pdparams"):
            path = path + ".pdparams"
        model.set_state_dict(checkpoint.load(path))
        return model

    def predict(self, mel):
        """reference :807-825: numpy mel (n_mels, T') -> numpy audio."""
        mel = torch.as_tensor(np.asarray(mel), dtype=torch.float32, device=self.device).unsqueeze(0)
        return self.infer(mel)[0].cpu().numpy()
