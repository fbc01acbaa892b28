"""TEST INFRASTRUCTURE ONLY.  Make the UNMODIFIED reference importable in the build container.

Used by `tests/golden/make_golden.py` (to generate golden vectors from the reference's own
code) and by the `not gpu` tests that pin the oracle restatement against the reference when
`/root/reference` is present.  Nothing here can run on the GPU box (no /root/reference there).

Shims, none of which touches reference code:
  * `easydict`, `termcolor`  -- tiny stand-ins (not installed in this image)
  * `os.popen('stty size')`  -- the detector modules call it at import time
    (faster_rcnn_coatt_transformer_sk.py:30, resnet_coatt_transformer_sk.py:33)
  * `model._C`               -- the reference's own CPU ops built by oracle/build_ref.py
"""
import importlib.util
import io
import os
import sys
import types

REF_ROOT = os.environ.get("AIT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "model"))


def load_ref_C():
    """Import the prebuilt oracle/_ref/aitref_C.so (building it first if sources exist)."""
    from . import build_ref
    path = build_ref.build()
    if "aitref_C" in sys.modules:
        return sys.modules["aitref_C"]
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("aitref_C", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["aitref_C"] = mod
    return mod


_installed = False


def install():
    """Put the reference's `lib/` on sys.path with the shims in place. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                d = dict(d or {}, **kw)
                for k, v in d.items():
                    setattr(self, k, v)

            def __setattr__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                dict.__setitem__(self, k, v)

            __setitem__ = __setattr__

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.colored = lambda s, *a, **k: s
        m.cprint = lambda s, *a, **k: print(s)
        sys.modules["termcolor"] = m

    _real_popen = os.popen

    def _popen(cmd, *a, **k):
        if isinstance(cmd, str) and cmd.strip().startswith("stty size"):
            return io.StringIO("40 120")
        return _real_popen(cmd, *a, **k)

    os.popen = _popen

    lib = os.path.join(REF_ROOT, "lib")
    for p in (REF_ROOT, lib):   # `lib.ops.utils` (resnet_coatt_transformer_sk.py:17) needs REF_ROOT too
        if p not in sys.path:
            sys.path.insert(0, p)
    import model  # the reference's package (lib/model/__init__.py)
    C = load_ref_C()
    sys.modules["model._C"] = C
    model._C = C
    _installed = True


def ref_transformer(**kw):
    """model.system.Models.Transformer -- what the detector imports
    (faster_rcnn_coatt_transformer_sk.py:27)."""
    install()
    from model.system.Models import Transformer
    args = dict(d_k=64, d_v=64, d_model=512, d_word_vec=512, d_inner=2048,
                n_position=64, n_layers=1, n_head=8, dropout=0.1)
    args.update(kw)
    return Transformer(**args)


def ref_sknet(channels=1024):
    install()
    from model.modules.blocks_coatt_transformer_sk import SKNet
    return SKNet(channels=channels)


def ref_layer4():
    """ResNet-50 layer4 exactly as `RCNN_top` is built (resnet_coatt_transformer_sk.py:416)."""
    install()
    import torch.nn as nn
    from model.faster_rcnn.resnet_coatt_transformer_sk import resnet50
    return nn.Sequential(resnet50().layer4)
