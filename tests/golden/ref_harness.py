"""Import harness for the UNMODIFIED reference (jasonkyuyim/se3_diffusion) — golden-vector generation only.

This module is used ONLY inside the build container, where ``/root/reference`` exists, by
``tests/golden/make_golden.py`` (and by ``tests/test_oracle_vs_reference.py`` when the reference is present).
Nothing on the product path, in the ``-m gpu`` tests, in ``smoke()`` or in ``bench.py`` imports it: the GPU
box has no ``/root/reference``.

The reference needs a handful of pure-Python packages that are not installed here (dm-tree, omegaconf,
ml_collections, biopython, ...).  None of them takes part in the arithmetic of the hot path, so they are
replaced by in-memory stub modules before the reference is imported.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import pickle
import sys
import types

REFERENCE_ROOT = os.environ.get("FRAMEDIFF_REFERENCE", "/root/reference")


class AttrDict(dict):
    """Minimal attribute-access dict standing in for an OmegaConf node."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


class _PermissiveModule(types.ModuleType):
    """Module whose unknown attributes resolve to an inert placeholder (e.g. `from tmtools import tm_align`)."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)

        def _placeholder(*a, **k):
            raise RuntimeError(f"stubbed symbol {self.__name__}.{item} called: not on the FrameDiff hot path")

        return _placeholder


def _stub(name, **attrs):
    m = _PermissiveModule(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, dict):
        return {k: _map_structure(fn, *[s[k] for s in structs]) for k in s0}
    if isinstance(s0, (list, tuple)):
        out = [_map_structure(fn, *[s[i] for s in structs]) for i in range(len(s0))]
        return type(s0)(out) if not hasattr(s0, "_fields") else type(s0)(*out)
    return fn(*structs)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


_installed = False


def install_stubs():
    """Register stub modules and put the reference tree on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # reference tree is read-only
    if "tree" not in sys.modules:
        _stub("tree", map_structure=_map_structure)
    if "omegaconf" not in sys.modules:
        class _OC:
            @staticmethod
            def to_container(x, **kw):
                return x

            @staticmethod
            def create(x=None, **kw):
                return to_attr(x or {})

            @staticmethod
            def merge(*xs):
                out = AttrDict()
                for x in xs:
                    out.update(x)
                return out

            @staticmethod
            def set_struct(*a, **k):
                return None

        _stub("omegaconf", OmegaConf=_OC, DictConfig=AttrDict)
    if "ml_collections" not in sys.modules:
        class _FieldReference:
            def __init__(self, v, field_type=None):
                self.v = v

            def get(self):
                return self.v

        class _ConfigDict(AttrDict):
            def __init__(self, d=None):
                super().__init__()
                for k, v in (d or {}).items():
                    if isinstance(v, _FieldReference):
                        v = v.get()
                    self[k] = _ConfigDict(v) if isinstance(v, dict) and not isinstance(v, _ConfigDict) else v

        _stub("ml_collections", ConfigDict=_ConfigDict, FieldReference=_FieldReference)
    if "Bio" not in sys.modules:
        bio = _stub("Bio")
        pdb = _stub("Bio.PDB", PDBParser=type("PDBParser", (), {}))
        chain = _stub("Bio.PDB.Chain", Chain=type("Chain", (), {}))
        bio.PDB = pdb
        pdb.Chain = chain
    for name in ("GPUtil", "mdtraj", "tmtools", "wandb"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "hydra" not in sys.modules:
        def _main(*a, **k):
            return lambda f: f

        hydra = _stub("hydra", main=_main)
        core = _stub("hydra.core")
        hc = _stub("hydra.core.hydra_config",
                   HydraConfig=type("HydraConfig", (), {"initialized": staticmethod(lambda: False)}))
        hydra.core = core
        core.hydra_config = hc
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # analysis/metrics.py imports openfold.np.relax.amber_minimize (needs pdbfixer/openmm): off the hot path.
    import openfold.np as _ofnp  # noqa: F401  (real package; only its `relax` sub-package is stubbed)
    relax = _stub("openfold.np.relax")
    relax.__path__ = []
    am = _stub("openfold.np.relax.amber_minimize")
    relax.amber_minimize = am
    _ofnp.relax = relax
    _installed = True


def load_experiment_class():
    """The reference's real `Experiment` class (for its unmodified inference_fn / loss_fn)."""
    install_stubs()
    from experiments import train_se3_diffusion as tsd
    return tsd.Experiment


def make_experiment(net, diffuser):
    """A light object carrying exactly the attributes Experiment.inference_fn touches, with the reference's own
    unbound methods bound to it (Experiment.__init__ wants a full hydra config + dataset; not needed here)."""
    exp_cls = load_experiment_class()
    model_conf, _ = default_conf()

    class _Exp:
        pass

    ex = _Exp()
    ex._model_conf = model_conf
    ex._data_conf = to_attr(dict(num_t=100, min_t=0.01))
    ex.model = ex._model = net
    ex.diffuser = ex._diffuser = diffuser
    for name in ("inference_fn", "_set_t_feats", "_self_conditioning"):
        setattr(ex, name, types.MethodType(getattr(exp_cls, name), ex))
    return ex


def default_conf(cache_dir="/tmp/framediff_igso3_cache", use_cached_score=False):
    """Attr-dict configs equal to config/base.yaml:25-67 (model + diffuser)."""
    diffuser = to_attr(dict(
        diffuse_trans=True, diffuse_rot=True,
        r3=dict(min_b=0.1, max_b=20.0, coordinate_scaling=0.1),
        so3=dict(num_omega=1000, num_sigma=1000, min_sigma=0.1, max_sigma=1.5, schedule="logarithmic",
                 cache_dir=cache_dir, use_cached_score=use_cached_score),
    ))
    model = to_attr(dict(
        node_embed_size=256, edge_embed_size=128, dropout=0.0,
        embed=dict(index_embed_size=32, aatype_embed_size=64, embed_self_conditioning=True,
                   num_bins=22, min_bin=1e-5, max_bin=20.0),
        ipa=dict(c_s=256, c_z=128, c_hidden=256, c_skip=64, no_heads=8, no_qk_points=8, no_v_points=12,
                 seq_tfmr_num_heads=4, seq_tfmr_num_layers=2, num_blocks=4, coordinate_scaling=0.1),
    ))
    return model, diffuser


class _PermissiveUnpickler(pickle.Unpickler):
    """Maps omegaconf.* (and anything else missing) to permissive dummies so weights/*.pth unpickle."""

    def find_class(self, module, name):
        if module.startswith("omegaconf") or module.startswith("typing") and name == "Any":
            class _Dummy(dict):
                def __init__(self, *a, **k):
                    dict.__init__(self)

                def __setstate__(self, state):
                    if isinstance(state, dict):
                        self.__dict__.update(state)

                def __reduce_ex__(self, p):  # pragma: no cover
                    return (dict, ())

            _Dummy.__name__ = name
            return _Dummy
        return super().find_class(module, name)


class _PickleModule:
    Unpickler = _PermissiveUnpickler
    __name__ = "pickle"

    @staticmethod
    def load(f, **kw):
        return _PermissiveUnpickler(f, **kw).load()


def load_reference_checkpoint(name="paper_weights.pth"):
    """state_dict of a shipped checkpoint with any leading 'module.' stripped."""
    import torch
    path = os.path.join(REFERENCE_ROOT, "weights", name)
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)
    sd = ckpt["model"]
    return {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}


def build_reference(state_dict=None, use_cached_score=False):
    """Returns (ScoreNetwork, SE3Diffuser) of the unmodified reference, CPU, eval mode."""
    install_stubs()
    import torch
    from data import se3_diffuser
    from model import score_network
    model_conf, diff_conf = default_conf(use_cached_score=use_cached_score)
    os.makedirs(diff_conf.so3.cache_dir, exist_ok=True)
    diffuser = se3_diffuser.SE3Diffuser(diff_conf)
    net = score_network.ScoreNetwork(model_conf, diffuser)
    if state_dict is not None:
        net.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()}, strict=True)
    net.eval()
    return net, diffuser
