"""ScoreNetwork — B200 mirror of the reference's model/score_network.py public API.

`ScoreNetwork(model_conf, diffuser)` is an nn.Module whose 282 parameters carry exactly the reference's names and shapes
(SURVEY.md Appendix A.6), so both shipped checkpoints load with strict=True, `.parameters()` / `.state_dict()` /
`.to(device)` / DataParallel wrappers keep working, and checkpoints written from it load back into the reference.
`forward(input_feats) -> dict` has the reference's keys, shapes and dtypes (model/score_network.py:170-215) but runs the
whole network in libframediff_b200.so.  The nn.Module is only the parameter container: no torch op takes part in the forward.

Training: with gradients enabled on trainable parameters, forward runs the training-mode CUDA forward (fd_train_forward: autograd
semantics of the sequence attention, activations kept on a tape) and returns tensors attached to a torch.autograd.Function whose
backward is the hand-written CUDA backward (fd_train_backward) — so the reference's `loss_fn(...)`, `loss.backward()`,
`torch.optim.Adam` and DDP wrappers (experiments/train_se3_diffusion.py:139-141,268-286,320-326) work unchanged.  The parameters
are views into one flat fp32 arena in state_dict order (the layout the kernels read and the gradient arena mirrors); the 10
parameters the reference never uses (linear_rbf, torsion_pred.linear_3) are not inputs of the Function and keep grad None, exactly
like under the reference's autograd (DDP find_unused_parameters=True).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .engine import FrameDiffEngine, arena_layout

_OUT_KEYS = ("psi", "rot_score", "trans_score", "rigids", "atom37", "atom14")


def _is_unused(name: str) -> bool:
    return ".linear_rbf." in name or name.startswith("score_model.torsion_pred.linear_3.")


class _ScoreNetworkFn(torch.autograd.Function):
    """ScoreNetwork.forward / backward on the CUDA training path.  Inputs after `net` and `feats` are the USED parameters (so that autograd
    — and DDP's reducer hooks — see them); outputs are the reference's six tensors."""

    @staticmethod
    def forward(ctx, net, feats, *params):
        eng = net._training_engine()
        out = eng.train_forward(feats)
        ctx.net = net
        ctx.dtypes = {k: out[k].dtype for k in _OUT_KEYS}
        return tuple(out[k] for k in _OUT_KEYS)

    @staticmethod
    def backward(ctx, *douts):
        net = ctx.net
        eng = net._training_engine()
        net._grad_flat.zero_()
        eng.train_backward({k: d for k, d in zip(_OUT_KEYS, douts)})
        return (None, None) + tuple(net._used_grad_views)

_TRUNC_STD = 0.87962566103423978   # std of a standard normal truncated to [-2, 2] (scipy.stats.truncnorm.std(-2, 2))


def _trunc_normal_(w: torch.Tensor, scale: float):
    fan_in = w.shape[1]
    std = math.sqrt(scale / max(1, fan_in)) / _TRUNC_STD
    return nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std)


class _Linear(nn.Linear):
    """Parameter holder with the reference's initialisers (model/ipa_pytorch.py:101-166): default = LeCun trunc-normal,
    relu = He trunc-normal, final = zeros; biases zero."""

    def __init__(self, in_dim, out_dim, init="default"):
        super().__init__(in_dim, out_dim, bias=True)
        with torch.no_grad():
            self.bias.fill_(0)
            if init == "default":
                _trunc_normal_(self.weight, 1.0)
            elif init == "relu":
                _trunc_normal_(self.weight, 2.0)
            elif init == "final":
                self.weight.fill_(0.0)
            else:
                raise ValueError("Invalid init string.")


class _IPA(nn.Module):
    def __init__(self, c):
        super().__init__()
        hc = c.c_hidden * c.no_heads
        self.linear_q = _Linear(c.c_s, hc)
        self.linear_kv = _Linear(c.c_s, 2 * hc)
        self.linear_q_points = _Linear(c.c_s, c.no_heads * c.no_qk_points * 3)
        self.linear_kv_points = _Linear(c.c_s, c.no_heads * (c.no_qk_points + c.no_v_points) * 3)
        self.linear_b = _Linear(c.c_z, c.no_heads)
        self.down_z = _Linear(c.c_z, c.c_z // 4)
        self.head_weights = nn.Parameter(torch.full((c.no_heads,), 0.541324854612918))
        self.linear_out = _Linear(c.no_heads * (c.c_z // 4 + c.c_hidden + c.no_v_points * 4), c.c_s, init="final")
        self.linear_rbf = _Linear(20, 1)          # unused by the reference as well; kept so checkpoints load strict


class _Transition(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.linear_1, self.linear_2, self.linear_3 = _Linear(c, c, "relu"), _Linear(c, c, "relu"), _Linear(c, c, "final")
        self.ln = nn.LayerNorm(c)


class _EdgeTransition(nn.Module):
    def __init__(self, node, edge):
        super().__init__()
        b = node // 2
        hid = 2 * b + edge
        self.initial_embed = _Linear(node, b, "relu")
        self.trunk = nn.Sequential(_Linear(hid, hid, "relu"), nn.ReLU(), _Linear(hid, hid, "relu"), nn.ReLU())
        self.final_layer = _Linear(hid, edge, "final")
        self.layer_norm = nn.LayerNorm(edge)


class _BackboneUpdate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.linear = _Linear(c, 6, "final")


class _Torsion(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.linear_1, self.linear_2, self.linear_3 = _Linear(c, c, "relu"), _Linear(c, c, "relu"), _Linear(c, c, "final")
        self.linear_final = _Linear(c, 2, "final")


class _IpaScore(nn.Module):
    def __init__(self, model_conf):
        super().__init__()
        c = model_conf.ipa
        self.trunk = nn.ModuleDict()
        for b in range(c.num_blocks):
            self.trunk[f"ipa_{b}"] = _IPA(c)
            self.trunk[f"ipa_ln_{b}"] = nn.LayerNorm(c.c_s)
            self.trunk[f"skip_embed_{b}"] = _Linear(model_conf.node_embed_size, c.c_skip, "final")
            d = c.c_s + c.c_skip
            layer = nn.TransformerEncoderLayer(d_model=d, nhead=c.seq_tfmr_num_heads, dim_feedforward=d, batch_first=True,
                                               dropout=0.0, norm_first=False)
            self.trunk[f"seq_tfmr_{b}"] = nn.TransformerEncoder(layer, c.seq_tfmr_num_layers, enable_nested_tensor=False)
            self.trunk[f"post_tfmr_{b}"] = _Linear(d, c.c_s, "final")
            self.trunk[f"node_transition_{b}"] = _Transition(c.c_s)
            self.trunk[f"bb_update_{b}"] = _BackboneUpdate(c.c_s)
            if b < c.num_blocks - 1:
                self.trunk[f"edge_transition_{b}"] = _EdgeTransition(c.c_s, model_conf.edge_embed_size)
        self.torsion_pred = _Torsion(c.c_s)


class _Embedder(nn.Module):
    def __init__(self, model_conf):
        super().__init__()
        e = model_conf.embed
        node_in = (e.index_embed_size + 1) + e.index_embed_size
        edge_in = (e.index_embed_size + 1) * 2 + e.index_embed_size + (e.num_bins if e.embed_self_conditioning else 0)
        n, z = model_conf.node_embed_size, model_conf.edge_embed_size
        self.node_embedder = nn.Sequential(nn.Linear(node_in, n), nn.ReLU(), nn.Linear(n, n), nn.ReLU(), nn.Linear(n, n), nn.LayerNorm(n))
        self.edge_embedder = nn.Sequential(nn.Linear(edge_in, z), nn.ReLU(), nn.Linear(z, z), nn.ReLU(), nn.Linear(z, z), nn.LayerNorm(z))


def _check_conf(model_conf):
    """The kernels are compiled for the shipped architecture (config/base.yaml:45-67, identical in both checkpoints)."""
    c, e = model_conf.ipa, model_conf.embed
    want = dict(node=256, edge=128, c_s=256, c_z=128, c_hidden=256, c_skip=64, heads=8, qk=8, v=12, th=4, tl=2, blocks=4, idx=32, bins=22)
    got = dict(node=model_conf.node_embed_size, edge=model_conf.edge_embed_size, c_s=c.c_s, c_z=c.c_z, c_hidden=c.c_hidden, c_skip=c.c_skip,
               heads=c.no_heads, qk=c.no_qk_points, v=c.no_v_points, th=c.seq_tfmr_num_heads, tl=c.seq_tfmr_num_layers, blocks=c.num_blocks,
               idx=e.index_embed_size, bins=e.num_bins)
    if want != got or not e.embed_self_conditioning or abs(float(c.coordinate_scaling) - 0.1) > 1e-12:
        raise ValueError(f"model_conf differs from the architecture the B200 kernels are compiled for: {got} vs {want}")


class ScoreNetwork(nn.Module):
    def __init__(self, model_conf, diffuser, precision: str = "bf16x3"):
        super().__init__()
        _check_conf(model_conf)
        self._model_conf = model_conf
        self.embedding_layer = _Embedder(model_conf)
        self.diffuser = diffuser
        self.score_model = _IpaScore(model_conf)
        self.precision = precision
        self._engine_obj = None
        self._loaded_version = None
        self._param_flat = None        # training: one flat fp32 arena the parameters are views of (engine.arena_layout order)
        self._grad_flat = None

    # ---- engine management ---------------------------------------------------------------------------------------------
    def _weights_key(self):
        """Changes whenever a parameter may have changed: in-place writes bump `_version`; `p.data = ...` / `.to()` / re-allocation change
        `data_ptr()`.  Writes through `p.data.copy_()` change neither — call `invalidate_weights()` after such utilities (EMA swaps)."""
        return tuple((p._version, p.data_ptr()) for p in self.parameters())

    def invalidate_weights(self):
        """Force the next forward to re-pack the device weight arena from the module's parameters."""
        self._loaded_version = None

    def _apply(self, fn, *a, **k):
        self._loaded_version = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._loaded_version = None
        return super().load_state_dict(*a, **k)

    def engine(self, device=None) -> FrameDiffEngine:
        dev = device if device is not None else next(self.parameters()).device
        if self._engine_obj is None or self._engine_obj.device != torch.device(dev):
            self._engine_obj = FrameDiffEngine(dev, self.precision)
            self._loaded_version = None
            if hasattr(self.diffuser, "bind_engine"):
                self.diffuser.bind_engine(self._engine_obj)
        if self._engine_obj.precision != self.precision:
            self._engine_obj.set_precision(self.precision)
        ver = self._weights_key()
        if ver != self._loaded_version:          # parameters changed (load_state_dict / optimiser step): repack the device arena
            self._engine_obj.load_weights(self.state_dict())
            self._loaded_version = ver
        return self._engine_obj

    # ---- training path -------------------------------------------------------------------------------------------------------
    def _flatten_parameters(self, dev):
        """(Re)points every parameter at its slice of one flat arena (values preserved).  Cheap check on every training forward, so
        `.to()`, `load_state_dict` into fresh tensors or external re-allocations are picked up."""
        lay, total = arena_layout()
        named = dict(self.named_parameters())
        flat = self._param_flat
        ok = flat is not None and flat.device == dev and all(
            named[n].data_ptr() == flat.data_ptr() + 4 * off and named[n].is_contiguous() for n, _, off in lay)
        if ok:
            return
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for n, shape, off in lay:
                view = flat[off:off + named[n].numel()].view(shape)
                view.copy_(named[n].detach().to(dev, torch.float32))
                named[n].data = view
        self._param_flat = flat
        self._grad_flat = torch.zeros_like(flat)
        gv = {n: self._grad_flat[off:off + named[n].numel()].view(shape) for n, shape, off in lay}
        self._used_names = [n for n, _, _ in lay if not _is_unused(n)]
        self._used_params = [named[n] for n in self._used_names]
        self._used_grad_views = [gv[n] for n in self._used_names]
        self._loaded_version = None

    def _training_engine(self) -> FrameDiffEngine:
        dev = self._param_flat.device
        if self._engine_obj is None or self._engine_obj.device != dev:
            self._engine_obj = FrameDiffEngine(dev, self.precision)
            self._loaded_version = None
            if hasattr(self.diffuser, "bind_engine"):
                self.diffuser.bind_engine(self._engine_obj)
        arenas = getattr(self._engine_obj, "_train_arenas", None)
        if arenas is None or arenas[0] is not self._param_flat:
            self._engine_obj.train_bind(self._param_flat, self._grad_flat)
        return self._engine_obj

    def forward(self, input_feats):
        dev = input_feats["rigids_t"].device
        if dev.type != "cuda":
            raise RuntimeError("ScoreNetwork (B200) has no CPU path: move the model and its inputs to a CUDA device")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad or self.training:
            if getattr(getattr(self.diffuser, "_so3_diffuser", None), "use_cached_score", False):
                raise ValueError("use_cached_score=True is an inference-only table look-up on the B200 path (piecewise constant in the angle); "
                                 "train with the shipped default use_cached_score=False")
            # torch's TransformerEncoder leaves its fused inference path whenever the module is in train() mode or autograd records; the
            # training-mode CUDA forward has those semantics (float key-padding mask added to the logits)
            self._flatten_parameters(torch.device(dev))
            if needs_grad:
                outs = _ScoreNetworkFn.apply(self, input_feats, *self._used_params)
                return dict(zip(_OUT_KEYS, outs))
            out = self._training_engine().train_forward(input_feats)
            return {k: out[k] for k in _OUT_KEYS}
        eng = self.engine(dev)
        out = eng.forward(input_feats)
        return {k: out[k] for k in _OUT_KEYS}
