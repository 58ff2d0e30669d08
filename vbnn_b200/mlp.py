"""The mlp.lua net object (reference mlp.lua:5-143) over libvbnn.so: buildModel, resetGradients,
sample, run, test, calc_lc, update -- same names and order as the reference so that a
main.lua-style loop drives it unchanged -- plus the fused per-minibatch entry points
(train_step / train_step_host) that keep minibatches, weights and noise on the GPU."""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from .config import opts_struct
from .context import Context, as_dev_f32, default_context
from .vblinear import Linear, VBLinear


class MLP:
    def __init__(self, opt=None, ctx: Context = None, max_batch=None):
        self.handle = None
        if opt is not None:
            self.buildModel(opt, ctx, max_batch)

    # ---- mlp.lua:7-60 ----
    def buildModel(self, opt, ctx: Context = None, max_batch=None):
        self.opt = opt
        self.ctx = ctx or default_context(seed=opt.get("seed", 3))
        sizes = [int(opt["input_size"])] + [int(h) for h in opt["hidden"]] + [len(opt["classes"])]
        self.sizes = sizes
        self.max_batch = int(max_batch or max(opt["batchSize"], opt.get("testBatchSize", 1)))
        arr = (C.c_int * len(sizes))(*sizes)
        o = opts_struct(opt)
        self.handle = C.c_void_p()
        L.check(L.lib().vbnn_mlp_create(self.ctx.handle, arr, len(sizes), 1 if opt.get("vb_output") else 0,
                                        self.max_batch, C.byref(o), C.byref(self.handle)))
        self.model = []
        self.vb_indices = []                                            # mlp.lua:9,15,23
        n = L.lib().vbnn_mlp_num_layers(self.handle)
        for k in range(n):
            h = C.c_void_p()
            L.check(L.lib().vbnn_mlp_layer(self.handle, k, C.byref(h)))
            vb = k < n - 1 or opt.get("vb_output")
            cls = VBLinear if vb else Linear
            self.model.append(cls(sizes[k], sizes[k + 1], opt, self.ctx, _borrowed=h))
            if vb:
                self.vb_indices.append(k)
        self._s = 0
        return self

    def init_params(self, seed=4, he_means=False):
        """means per VBLinear.lua:22-29 (or He-scaled), Linear weights N(0, 2/fan_in)
        and zero biases per mlp.lua:47-55, drawn on the device."""
        L.check(L.lib().vbnn_mlp_init_params(self.handle, C.c_uint64(seed), 1 if he_means else 0))

    def __del__(self):
        try:
            if self.handle:
                for m in self.model:
                    m.handle = None
                L.lib().vbnn_mlp_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _io(self, inputs, targets):
        x = as_dev_f32(inputs, self.ctx.device)
        x = x.reshape(x.shape[0], -1)                                   # nn.Reshape (mlp.lua:12)
        t = as_dev_f32(targets, self.ctx.device).reshape(-1)
        if x.shape[1] != self.sizes[0] or t.shape[0] != x.shape[0]:
            raise L.VbnnError(L.E_INVALID, f"inputs must be [N x {self.sizes[0]}] with N targets")
        return x, t

    # ---- mlp.lua:62-84 ----
    def resetGradients(self):
        L.check(L.lib().vbnn_mlp_reset_gradients(self.handle))
        self._s = 0

    def sample(self, sample_idx=None):
        if sample_idx is not None:
            self._s = int(sample_idx)
        L.check(L.lib().vbnn_mlp_sample(self.handle, self._s))
        self._cur = self._s
        self._s += 1

    def run(self, inputs, targets):
        x, t = self._io(inputs, targets)
        err, acc = C.c_float(), C.c_float()
        s = getattr(self, "_cur", 0)
        if self.opt.get("reparam", "weight") == "local":
            s = self._s
            self._s += 1
        L.check(L.lib().vbnn_mlp_run(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                     x.shape[0], s, C.byref(err), C.byref(acc)))
        return err.value, acc.value

    # ---- mlp.lua:86-107 ----
    def test(self, input, target):
        x, t = self._io(input, target)
        err, acc = C.c_float(), C.c_float()
        ns = 0 if self.opt.get("quicktest") else int(self.opt["testSamples"])
        L.check(L.lib().vbnn_mlp_test(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                      x.shape[0], ns, C.byref(err), C.byref(acc)))
        return err.value, acc.value

    def calc_lc(self, opt=None):                                        # mlp.lua:109-115
        v = C.c_float()
        L.check(L.lib().vbnn_mlp_calc_lc(self.handle, C.byref(v)))
        return v.value

    def update(self, opt=None):                                         # mlp.lua:117-142
        opt = opt or self.opt
        from . import logger
        if opt.get("log") and logger.Log is not None and self.ctx.nranks == 1:
            # with logging on the reference syncs >= 13 scalars per layer anyway (VBLinear.lua:150-163): go
            # layer by layer so every VBLinear reports its 14 diagnostics -- output-layer SGD first
            # (mlp.lua:120-123), then the VB layers in index order (:138-140); all layers log to the same
            # ids, interleaved, exactly as the reference does
            if self.model[-1].kind == L.KIND_LINEAR:
                L.check(L.lib().vbnn_layer_update(self.model[-1].handle, None))
            for k in self.vb_indices:
                self.model[k].update(opt)
            self.ctx.set_step(self.ctx.get_step() + 1)
            return
        L.check(L.lib().vbnn_mlp_update(self.handle))

    # ---- fused minibatch: main.lua:28-40 in one call ----
    def train_step(self, inputs, targets, sync=True):
        """inputs/targets already on the device.  Returns (error, accuracy) averaged over the S
        samples as main.lua:38-39 (or None when sync=False)."""
        import torch
        x, t = self._io(inputs, targets)
        if not hasattr(self, "_res"):
            self._res = torch.zeros(2, dtype=torch.float32, device=x.device)
        L.check(L.lib().vbnn_mlp_step(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                      x.shape[0], C.c_void_p(self._res.data_ptr())))
        if not sync:
            return None
        r = self._res.cpu()
        return float(r[0]), float(r[1])

    def train_step_host(self, inputs_host, targets_host):
        """HOST buffers in, host scalars out: H2D and D2H inside the call (main.lua:23-24)."""
        x, t = self._host_io(inputs_host, targets_host)
        err, acc = C.c_float(), C.c_float()
        L.check(L.lib().vbnn_mlp_step_host(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                           x.shape[0], C.byref(err), C.byref(acc)))
        return err.value, acc.value

    def _host_io(self, inputs_host, targets_host):
        import torch
        x = torch.as_tensor(inputs_host, dtype=torch.float32)
        t = torch.as_tensor(targets_host, dtype=torch.float32)
        if x.is_cuda or t.is_cuda:
            raise L.VbnnError(L.E_INVALID, "train_step_host expects host tensors")
        x = x.reshape(x.shape[0], -1).contiguous()
        return x, t.reshape(-1).contiguous()

    def submit_host(self, inputs_host, targets_host):
        x, t = self._host_io(inputs_host, targets_host)
        self._keep = getattr(self, "_keep", [])
        self._keep.append((x, t))
        self._keep = self._keep[-4:]
        L.check(L.lib().vbnn_mlp_submit_host(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                             x.shape[0]))

    def submit_host_u8(self, pixels_host, targets_host, mean, std):
        """The dataset's native bytes (MNIST pixels before data.lua:25 u.normalize): uint8 [N x input_size] host
        tensor; (x - mean) / std is applied on the device while staging the GEMM operand."""
        import torch
        x = torch.as_tensor(pixels_host)
        if x.dtype != torch.uint8 or x.is_cuda:
            raise L.VbnnError(L.E_INVALID, "submit_host_u8 expects a uint8 host tensor")
        x = x.reshape(x.shape[0], -1).contiguous()
        t = torch.as_tensor(targets_host, dtype=torch.float32).reshape(-1).contiguous()
        self._keep = getattr(self, "_keep", [])
        self._keep.append((x, t))
        self._keep = self._keep[-4:]
        L.check(L.lib().vbnn_mlp_submit_host_u8(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                                x.shape[0], C.c_float(mean), C.c_float(1.0 / std)))

    def join_streams(self):
        L.check(L.lib().vbnn_mlp_join_streams(self.handle))

    def collect(self):
        err, acc = C.c_float(), C.c_float()
        L.check(L.lib().vbnn_mlp_collect(self.handle, C.byref(err), C.byref(acc)))
        return err.value, acc.value

    def outputs(self, sample_idx=0, n=None):
        """LogSoftMax output of the last forward (self.model.output in the reference)."""
        import torch
        n = n or self._last_n()
        out = torch.empty(n, self.sizes[-1], dtype=torch.float32)
        L.check(L.lib().vbnn_mlp_get_outputs(self.handle, sample_idx, C.c_void_p(out.data_ptr())))
        return out

    def _last_n(self):
        return self.max_batch

    # ---- checkpoint / resume (reference: u.safe_save(net) main.lua:181, utils.lua:73-80; resume
    # main.lua:146-148).  The reference torch.save()s the whole object graph; here the device state
    # is exported as a flat dict of CPU tensors (means, lvars, bias, Adam m/v/t per layer, the output
    # layer, the Philox step counter) that torch.save / numpy can persist.
    def state_dict(self):
        sd = {"sizes": list(self.sizes), "step": self.ctx.get_step()}
        names = [("means", L.BUF_MEANS), ("lvars", L.BUF_LVARS), ("bias", L.BUF_BIAS), ("m_mu", L.BUF_ADAM_M_MU),
                 ("v_mu", L.BUF_ADAM_V_MU), ("m_var", L.BUF_ADAM_M_VAR), ("v_var", L.BUF_ADAM_V_VAR)]
        for k, m in enumerate(self.model):
            if m.kind == L.KIND_VB:
                for n, b in names:
                    sd[f"{k}.{n}"] = m.get(b)
                if self.opt.get("strict_reference", True):
                    sd[f"{k}.stdv"] = m.get(L.BUF_STDV)
                    sd[f"{k}.mu_sqe"] = m.get(L.BUF_MU_SQE)
            else:
                sd[f"{k}.weight"] = m.get(L.BUF_WEIGHT)
                sd[f"{k}.bias"] = m.get(L.BUF_BIAS)
            sd[f"{k}.t"] = m.t
        return sd

    def load_state_dict(self, sd):
        if list(sd["sizes"]) != list(self.sizes):
            raise L.VbnnError(L.E_INVALID, f"checkpoint is for sizes {sd['sizes']}, net has {self.sizes}")
        codes = dict(means=L.BUF_MEANS, lvars=L.BUF_LVARS, bias=L.BUF_BIAS, m_mu=L.BUF_ADAM_M_MU, v_mu=L.BUF_ADAM_V_MU,
                     m_var=L.BUF_ADAM_M_VAR, v_var=L.BUF_ADAM_V_VAR, weight=L.BUF_WEIGHT, stdv=L.BUF_STDV,
                     mu_sqe=L.BUF_MU_SQE)
        for k, m in enumerate(self.model):
            L.check(L.lib().vbnn_layer_set_t(m.handle, int(sd[f"{k}.t"])))
            for n, b in codes.items():
                if f"{k}.{n}" in sd:
                    m.set(b, sd[f"{k}.{n}"])
            if m.kind == L.KIND_VB and f"{k}.stdv" not in sd:
                m.compute_prior()
        self.ctx.set_step(int(sd["step"]))

    # ---- data parallel over NVLink peer memory (new; include/vbnn.h "peer mode") ----
    def enable_peer(self, all_gather_bytes):
        """all_gather_bytes(blob: bytes) -> list[bytes] in rank order (e.g. over torch.distributed).
        After this, train_step / submit_host exchange gradients through the fused dW-epilogue
        reduce-scatter + sharded update + copy-engine all-gather instead of an NCCL allreduce."""
        n = C.c_size_t()
        L.check(L.lib().vbnn_mlp_peer_export(self.handle, None, 0, C.byref(n)))
        buf = (C.c_char * n.value)()
        L.check(L.lib().vbnn_mlp_peer_export(self.handle, buf, n.value, C.byref(n)))
        blobs = all_gather_bytes(bytes(buf))
        if len(blobs) != self.ctx.nranks or any(len(b) != n.value for b in blobs):
            raise L.VbnnError(L.E_INVALID, "enable_peer: all_gather_bytes must return one blob per rank")
        joined = b"".join(blobs)
        L.check(L.lib().vbnn_mlp_peer_import(self.handle, joined, n.value))

    @property
    def peer_active(self):
        return bool(L.lib().vbnn_mlp_peer_active(self.handle))

    def sync_replicas(self):
        """Collective: call on every rank between two host barriers (see vbnn_mlp_sync_replicas)."""
        L.check(L.lib().vbnn_mlp_sync_replicas(self.handle))

    def launch_count(self):
        v = C.c_longlong()
        L.check(L.lib().vbnn_mlp_launch_count(self.handle, C.byref(v)))
        return v.value

    def grad_arena(self):
        import torch
        from .context import DevView
        p, n = C.c_void_p(), C.c_size_t()
        L.check(L.lib().vbnn_mlp_grad_arena(self.handle, C.byref(p), C.byref(n)))
        return torch.as_tensor(DevView(p.value, (n.value,)), device=f"cuda:{self.ctx.device}")
