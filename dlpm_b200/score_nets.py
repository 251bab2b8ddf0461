"""Score networks of the hot path with the reference's constructors and ``state_dict`` key names.

  * ``MLPModel(p)``      -- mirror of ``dlpm/models/Model.py:17-211`` (2-D configs)
  * ``UNetModel(...)``   -- mirror of ``dlpm/models/unet.py:276-492`` (image configs)

These classes are *parameter containers*: they register exactly the reference's parameter names
and shapes (so ``load_state_dict`` of a reference checkpoint works, SURVEY.md section 5 "checkpoint"),
but ``forward`` never runs PyTorch ops -- it packs the weights once into the layout the CUDA
engine wants and calls the C ABI (``dlpm_b200_mlp_forward`` / ``dlpm_b200_unet_forward``).
There is no eager / CPU fallback.
"""
import math

import torch
import torch.nn as nn

from . import _lib


def _named(**mods):
    """An nn.Module whose children carry the given (possibly numeric) names."""
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(k, v)
    return m


class _Indexable(nn.Module):
    """Sequential-like container with explicit (numeric) child names; ``m[i]`` -> child ``str(i)``."""

    def __getitem__(self, idx):
        return getattr(self, str(idx))


def _seq(*pairs):
    m = _Indexable()
    for name, mod in pairs:
        m.add_module(str(name), mod)
    return m


class _PackedCache:
    """Caches a packed device copy of the parameters, invalidated by parameter version bumps."""

    def __init__(self):
        self.key = None
        self.value = None

    def get(self, module, build):
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in module.parameters())
        if key != self.key:
            self.value = build()
            self.key = key
        return self.value


# ------------------------------------------------------------------------------------------------
# MLP (2-D configs)
# ------------------------------------------------------------------------------------------------
class MLPModel(nn.Module):
    """Time-conditioned residual MLP; constructor reads the same ``p`` dict as Model.py:24-42.

    Only the configuration that can run in the reference is supported (no_a=True, learnable time
    embedding, LayerNorm, skip connections, learn_variance=False)."""

    def __init__(self, p):
        super().__init__()
        pm = p["model"]
        self.nfeatures = p["data"]["nfeatures"]
        self.nunits = pm["nunits"]
        self.nblocks = pm["nblocks"]
        self.time_emb_size = pm["time_emb_size"]
        self.device_name = p.get("device", "cuda")
        assert pm["no_a"] and p[p["method"]]["isotropic"], \
            "Need to reimplement architecture if model takes non-isotropic a_t as input."
        if pm["time_emb_type"] != "learnable" or pm.get("a_pos_emb", False) or pm.get("learn_variance", False) \
                or not pm.get("group_norm", True) or not pm.get("skip_connection", True):
            raise NotImplementedError("dlpm_b200 MLPModel supports the shipped 2d_data.yml architecture only "
                                      "(learnable time embedding, LayerNorm, skip connections, fixed variance)")
        U, E, F = self.nunits, self.time_emb_size, self.nfeatures
        # registration mirrors the reference so state_dict keys match (incl. its aliased entries)
        self.group_norm_in = nn.LayerNorm([U])
        self.time_emb = nn.Linear(1, E)
        self.time_mlp = _seq((0, self.time_emb), (2, nn.Linear(E, E)))
        self.linear_in = nn.Linear(F, U)
        self.inblock = _seq((0, self.linear_in), (1, self.group_norm_in))
        self.midblocks = nn.ModuleList([self._block(U, E) for _ in range(self.nblocks)])
        self.outblocks_mean = nn.ModuleList([self._block(U, E), nn.Linear(U, F)])
        self._cache = _PackedCache()

    @staticmethod
    def _block(U, E):
        gn1, gn2 = nn.LayerNorm([U]), nn.LayerNorm([U])
        return _named(group_norm1=gn1, group_norm2=gn2, mlp_1=_seq((1, nn.Linear(U, U)), (2, gn1)),
                      t_proj=_seq((1, nn.Linear(E, U))), mlp_2=_seq((1, nn.Linear(U, U)), (2, gn2)))

    @property
    def nblocks_total(self):
        return self.nblocks + 1

    def packed_weights(self):
        """Flat fp32 device buffer in the layout of ``include/dlpm_b200.h`` (K4) / ``csrc/mlp.cu``."""
        def build():
            f = lambda t: t.detach().float().reshape(-1)
            T = lambda lin: lin.weight.detach().float().t().contiguous().reshape(-1)  # in-major W^T
            blocks = list(self.midblocks) + [self.outblocks_mean[0]]
            parts = [f(self.time_mlp[0].weight), f(self.time_mlp[0].bias), T(self.time_mlp[2]), f(self.time_mlp[2].bias)]
            for b in blocks:
                parts += [T(b.t_proj[1]), f(b.t_proj[1].bias)]
            parts += [T(self.linear_in), f(self.linear_in.bias), f(self.group_norm_in.weight), f(self.group_norm_in.bias)]
            for b in blocks:
                parts += [T(b.mlp_1[1]), f(b.mlp_1[1].bias), f(b.mlp_1[2].weight), f(b.mlp_1[2].bias),
                          T(b.mlp_2[1]), f(b.mlp_2[1].bias), f(b.mlp_2[2].weight), f(b.mlp_2[2].bias)]
            out = self.outblocks_mean[1]
            pad = torch.zeros(((self.nfeatures + 3) // 4) * 4 - self.nfeatures, device=out.bias.device)
            parts += [f(out.weight), f(out.bias), pad]
            return torch.cat(parts).contiguous()
        return self._cache.get(self, build)

    def forward(self, x, timestep):
        """x: (B, 1, nfeatures) CUDA fp32; timestep: (B,) already scaled (Model.py:148)."""
        w = self.packed_weights()
        _lib.require_cuda(w.device)
        B = x.shape[0]
        xin = x.to(w.device, torch.float32).reshape(B, self.nfeatures).contiguous()
        t = timestep.to(w.device, torch.float32).reshape(B).contiguous()
        out = torch.empty_like(xin)
        with torch.cuda.device(w.device):
            _lib.call("dlpm_b200_mlp_forward", _lib.ptr(out), _lib.ptr(xin), _lib.ptr(t), _lib.ptr(w), B,
                      self.nfeatures, self.nunits, self.time_emb_size, self.nblocks_total, _lib.stream_ptr())
        return out.reshape(x.shape)

    native_kind = "mlp"


def _mlp_from_reference(ref, device):
    """Ingest a reference ``MLPModel`` instance (dlpm/models/Model.py) by hyper-parameters + state_dict."""
    p = {"data": {"nfeatures": ref.nfeatures}, "method": "dlpm", "dlpm": {"isotropic": True}, "device": str(device),
         "model": dict(use_a_t=ref.use_a_t, no_a=ref.no_a, a_pos_emb=ref.a_pos_emb, a_emb_size=ref.a_emb_size,
                       time_emb_type=ref.time_emb_type, time_emb_size=ref.time_emb_size, nblocks=ref.nblocks,
                       nunits=ref.nunits, skip_connection=ref.skip_connection, group_norm=ref.group_norm,
                       dropout_rate=ref.dropout_rate, learn_variance=ref.learn_variance)}
    m = MLPModel(p)
    m.load_state_dict(ref.state_dict(), strict=True)
    return m.to(device).eval()


def as_native(model, device):
    """Return a module whose forward runs on the CUDA engine when ``model`` is one of the hot-path
    score nets (this package's classes, or the reference's ``MLPModel`` / ``UNetModel`` which are
    ingested through their ``state_dict``); any other ``nn.Module`` is returned unchanged and is simply
    called on the device by the sampling loop."""
    if getattr(model, "native_kind", None) is not None:
        return model
    cached = getattr(model, "_dlpm_b200_native", None)
    if cached is not None:
        return cached
    name = type(model).__name__
    native = None
    if name == "MLPModel" and hasattr(model, "midblocks") and hasattr(model, "outblocks_mean"):
        native = _mlp_from_reference(model, device)
    elif name == "UNetModel" and hasattr(model, "input_blocks") and hasattr(model, "output_blocks"):
        native = _unet_from_reference(model, device)
    if native is None:
        return model
    try:
        object.__setattr__(model, "_dlpm_b200_native", native)
    except Exception:
        pass
    return native


def _unet_from_reference(ref, device):
    raise NotImplementedError("UNet engine not built yet")
