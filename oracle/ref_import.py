"""Import the UNMODIFIED reference (``/root/reference``) for pinning the oracle.  TEST INFRASTRUCTURE.

The reference is pure Python/PyTorch; its hot-path modules import with torch only, except
``models/mvsformer_model.py`` which pulls ``timm`` (models/gvt.py:6-7) and ``omegaconf``
(utils.py:13), both absent here and both OFF the hot path.  We register inert stand-ins for
those two packages in ``sys.modules`` and import the reference files
from where they lie.  Nothing is copied into this repository.

``/root/reference`` exists only in the build container, never on the GPU box; there the staged, unmodified
copy ``oracle/_ref/`` (git-ignored; made by ``oracle/build_ref.py``, travels with the gpurun snapshot) is
used.  Callers must use ``reference_available()`` and skip otherwise.
"""
import importlib
import os
import sys
import types
import warnings

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    env = os.environ.get("MVS_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(root, "models", "mvsformer_model.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "mvsformer_model.py"))


def _install_stubs():
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        vit = types.ModuleType("timm.models.vision_transformer")
        registry = types.ModuleType("timm.models.registry")

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                return x

        class _Attn(nn.Module):
            def __init__(self, dim, num_heads=8, qkv_bias=False, **kw):
                super().__init__()
                self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
                self.proj = nn.Linear(dim, dim)

        class _Mlp(nn.Module):
            def __init__(self, dim, hidden, act_layer=nn.GELU, drop=0.0):
                super().__init__()
                self.fc1 = nn.Linear(dim, hidden)
                self.act = act_layer()
                self.fc2 = nn.Linear(hidden, dim)
                self.drop = nn.Dropout(drop)

            def forward(self, x):
                return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))

        class Block(nn.Module):
            def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0,
                         attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, **kw):
                super().__init__()
                self.norm1 = norm_layer(dim)
                self.attn = _Attn(dim, num_heads, qkv_bias)
                self.drop_path = DropPath(drop_path)
                self.norm2 = norm_layer(dim)
                self.mlp = _Mlp(dim, int(dim * mlp_ratio), act_layer, drop)

        layers.DropPath = DropPath
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        layers.trunc_normal_ = nn.init.trunc_normal_
        vit.Block = Block
        vit._cfg = lambda **kw: dict(kw)
        registry.register_model = lambda fn: fn
        timm.models = timm_models
        timm_models.layers = layers
        timm_models.vision_transformer = vit
        timm_models.registry = registry
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": layers,
                            "timm.models.vision_transformer": vit, "timm.models.registry": registry})
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        oc.OmegaConf = type("OmegaConf", (), {})
        sys.modules["omegaconf"] = oc


_CACHE = {}


def load_reference():
    """Returns a namespace with the reference's ``warping``, ``module`` and ``mvsformer_model`` modules."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's package is called ``models``; make sure nothing of ours shadows it
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        if not getattr(sys.modules[name], "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        warping = importlib.import_module("models.warping")
        module = importlib.import_module("models.module")
        model = importlib.import_module("models.mvsformer_model")
    ns = types.SimpleNamespace(warping=warping, module=module, mvsformer_model=model)
    _CACHE["ns"] = ns
    return ns
